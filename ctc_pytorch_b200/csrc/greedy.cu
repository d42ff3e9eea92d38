// K8 — frame-wise arg-max and CTC greedy collapse.
//
// Replaces torch.max(out, -1) at timit/steps/train_ctc.py:51, the Python collapse loop of
// CTC_Model.compute_wer (timit/models/model_ctc.py:187-202) and GreedyDecoder.decode's
// arg-max + remove-repeat + drop-blank (timit/utils/ctcDecoder.py:162-166, 79-92).
// Integer work: results are bit-identical to the reference (ties resolve to the first index, a
// label is kept iff it is not blank and differs from the label of the previous *frame*).
#include "common.cuh"
#include "ctcb200.h"

namespace ctcb200 {
namespace {

// One warp per (t, n) row of the [T, N, C] log-prob tensor; writes idx[n][t].
__global__ void __launch_bounds__(256)
argmax_rows_kernel(const float* __restrict__ lp, int* __restrict__ idx, float* __restrict__ maxv, int T, int N,
                   int C) {
    const int warps_per_block = blockDim.x >> 5;
    const int lane = threadIdx.x & 31;
    const long long rows = static_cast<long long>(T) * N;
    for (long long row = static_cast<long long>(blockIdx.x) * warps_per_block + (threadIdx.x >> 5); row < rows;
         row += static_cast<long long>(gridDim.x) * warps_per_block) {
        const float* p = lp + row * C;
        float best = -INFINITY;
        int bi = -1;  // -1 = this lane saw no element
        for (int c = lane; c < C; c += 32) {
            float v = __ldg(p + c);
            // first maximum wins; a NaN beats every number (torch.max semantics), first NaN wins
            if (bi < 0 || v > best || (v != v && best == best)) { best = v; bi = c; }
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            float ov = __shfl_xor_sync(0xffffffffu, best, o);
            int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            bool other_nan = ov != ov, mine_nan = best != best;
            bool take;
            if (oi < 0) take = false;
            else if (bi < 0) take = true;
            else if (mine_nan || other_nan) take = other_nan && (!mine_nan || oi < bi);
            else take = (ov > best) || (ov == best && oi < bi);
            if (take) { best = ov; bi = oi; }
        }
        if (lane == 0) {
            int t = static_cast<int>(row / N), n = static_cast<int>(row % N);
            idx[static_cast<size_t>(n) * T + t] = bi;
            if (maxv) maxv[static_cast<size_t>(n) * T + t] = best;
        }
    }
}

// One block per utterance: stream compaction of the kept frames.
__global__ void __launch_bounds__(256)
collapse_kernel(const int* __restrict__ idx, const int64_t* __restrict__ lengths, int* __restrict__ out,
                int* __restrict__ out_len, int T, int blank) {
    __shared__ int warp_tot[8];
    __shared__ int base;
    const int n = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int len = static_cast<int>(lengths[n]);
    if (len > T) len = T;
    if (len < 0) len = 0;
    const int* in = idx + static_cast<size_t>(n) * T;
    int* o = out + static_cast<size_t>(n) * T;
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    for (int t0 = 0; t0 < len; t0 += blockDim.x) {
        int t = t0 + threadIdx.x;
        int v = (t < len) ? in[t] : blank;
        bool keep = (t < len) && v != blank && (t == 0 || v != in[t - 1]);
        unsigned m = __ballot_sync(0xffffffffu, keep);
        int pre = __popc(m & ((1u << lane) - 1));
        if (lane == 0) warp_tot[warp] = __popc(m);
        __syncthreads();
        int off = base;
        for (int w = 0; w < warp; ++w) off += warp_tot[w];
        if (keep) o[off + pre] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
            int tot = 0;
            for (int w = 0; w < (blockDim.x >> 5); ++w) tot += warp_tot[w];
            base += tot;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) out_len[n] = base;
}

}  // namespace
}  // namespace ctcb200

using namespace ctcb200;

extern "C" CTCB200_API int ctcb200_argmax(const float* log_probs, int T, int N, int C, int* idx_nt, float* max_nt,
                              ctcb200_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    CTCB_REQUIRE(T > 0 && N > 0 && C > 0, "argmax: empty shape T=%d N=%d C=%d", T, N, C);
    long long rows = static_cast<long long>(T) * N;
    int blocks = static_cast<int>((rows + 7) / 8);
    int cap = device_sm_count() * 8;
    if (blocks > cap) blocks = cap;
    argmax_rows_kernel<<<blocks, 256, 0, stream>>>(log_probs, idx_nt, max_nt, T, N, C);
    CTCB_LAUNCH_CHECK();
    return OK;
}

extern "C" CTCB200_API int ctcb200_greedy_decode(const float* log_probs, const int64_t* lengths, int T, int N, int C, int blank,
                                     int* idx_nt, int* labels_nt, int* label_lengths, ctcb200_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    int rc = ctcb200_argmax(log_probs, T, N, C, idx_nt, nullptr, stream);
    if (rc != OK) return rc;
    collapse_kernel<<<N, 256, 0, stream>>>(idx_nt, lengths, labels_nt, label_lengths, T, blank);
    CTCB_LAUNCH_CHECK();
    return OK;
}
