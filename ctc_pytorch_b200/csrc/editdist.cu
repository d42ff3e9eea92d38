// K10 — batched Levenshtein distance on the device: replaces the per-utterance Python DP of
// Decoder._edit_distance (timit/utils/ctcDecoder.py:131-149) and the editdistance.eval call inside
// CTC_Model.compute_wer (timit/models/model_ctc.py:187-202) for integer label sequences.
//
// One warp per pair (hypothesis a_n, reference b_n). The DP row over the reference positions j = 0..Lb is blocked over
// the lanes (KS consecutive j per lane); the hypothesis tokens are consumed one per step. The in-row dependency
// d[j] = min(d[j-1] + 1, .) is a prefix minimum: with u[j] = min(prev[j] + 1, prev[j-1] + (a_i != b_j)) the new row is
// d[j] = j + min_{k <= j} (u[k] - k), evaluated with a lane-local scan plus a 5-step warp min-scan. Unit costs for
// insertion, deletion and substitution — the reference's DP — so results are exact integers.
#include "common.cuh"
#include "ctcb200.h"

namespace ctcb200 {
namespace {

template <int KS>
__global__ void __launch_bounds__(128)
edit_distance_kernel(const int* __restrict__ a, long long a_stride, const int* __restrict__ a_len,
                     const long long* __restrict__ b, long long b_stride, const long long* __restrict__ b_len, int N,
                     int* __restrict__ dist) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = blockIdx.x * (blockDim.x >> 5) + warp;
    if (n >= N) return;
    const int La = a_len[n];
    const int Lb = static_cast<int>(b_len[n]);
    const int* an = a + static_cast<long long>(n) * a_stride;
    const long long* bn = b + static_cast<long long>(n) * b_stride;
    constexpr int BIG = 1 << 29;
    // this lane owns row entries j = lane*KS + 1 .. lane*KS + KS (entry 0 is the row index itself)
    int bj[KS], prev[KS];
#pragma unroll
    for (int k = 0; k < KS; ++k) {
        const int j = lane * KS + k + 1;
        bj[k] = (j <= Lb) ? static_cast<int>(bn[j - 1]) : -1;
        prev[k] = (j <= Lb) ? j : BIG;
    }
    for (int i = 1; i <= La; ++i) {
        const int ai = an[i - 1];
        // prev[j-1] of this lane's first entry comes from the lane below (or is the border value i-1 for j = 1)
        int left = __shfl_up_sync(0xffffffffu, prev[KS - 1], 1);
        if (lane == 0) left = i - 1;
        int u[KS];
        int run = BIG;  // lane-local prefix minimum of u[k] - j
#pragma unroll
        for (int k = 0; k < KS; ++k) {
            const int j = lane * KS + k + 1;
            const int diag = (k == 0) ? left : prev[k - 1];
            const int cand = min(prev[k] + 1, diag + (ai != bj[k] ? 1 : 0));
            run = min(run, (j <= Lb ? cand : BIG) - j);
            u[k] = run;
        }
        // exclusive warp scan of the lane totals; the border entry d[i][0] = i contributes i - 0
        int carry = u[KS - 1];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, carry, o);
            if (lane >= o) carry = min(carry, v);
        }
        int excl = __shfl_up_sync(0xffffffffu, carry, 1);
        if (lane == 0) excl = BIG;
        excl = min(excl, i);
#pragma unroll
        for (int k = 0; k < KS; ++k) {
            const int j = lane * KS + k + 1;
            prev[k] = (j <= Lb) ? min(u[k], excl) + j : BIG;
        }
    }
    // d[La][Lb]
    int res = (Lb == 0) ? La : BIG;
#pragma unroll
    for (int k = 0; k < KS; ++k)
        if (lane * KS + k + 1 == Lb) res = prev[k];
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) res = min(res, __shfl_xor_sync(0xffffffffu, res, o));
    if (lane == 0) dist[n] = res;
}

}  // namespace
}  // namespace ctcb200

using namespace ctcb200;

extern "C" CTCB200_API int ctcb200_edit_distance(const int32_t* a, int64_t a_stride, const int32_t* a_len, const int64_t* b,
                                                 int64_t b_stride, const int64_t* b_len, int N, int max_b_len, int32_t* dist,
                                                 ctcb200_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    CTCB_REQUIRE(N > 0, "edit_distance: empty batch N=%d", N);
    CTCB_REQUIRE(max_b_len >= 0 && max_b_len <= 1024, "edit_distance: reference length %d exceeds the supported maximum 1024", max_b_len);
    const dim3 grid((N + 3) / 4), block(128);
    const long long* bp = reinterpret_cast<const long long*>(b);
    const long long* blp = reinterpret_cast<const long long*>(b_len);
#define LAUNCH_ED(KS) edit_distance_kernel<KS><<<grid, block, 0, stream>>>(a, a_stride, a_len, bp, b_stride, blp, N, dist)
    if (max_b_len <= 32) LAUNCH_ED(1);
    else if (max_b_len <= 64) LAUNCH_ED(2);
    else if (max_b_len <= 128) LAUNCH_ED(4);
    else if (max_b_len <= 256) LAUNCH_ED(8);
    else if (max_b_len <= 512) LAUNCH_ED(16);
    else LAUNCH_ED(32);
#undef LAUNCH_ED
    CTCB_LAUNCH_CHECK();
    return OK;
}
