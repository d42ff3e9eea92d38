// K11 — device-side batch assembly: the per-utterance feature pipeline of SpeechDataset.__getitem__
// (timit/utils/data_loader.py:103-110: make_context -> skip_feat -> zero rows up to a multiple of n_downsample;
// timit/utils/tools.py:66-86) and the padding of create_input (data_loader.py:119-140) as one gather kernel over the
// raw, concatenated utterance features already resident on the device.
//
//   ctx[t, c*F + f] = feat[clamp(t + c - left, 0, L-1), f]          c = 0 .. left+right   (edge frames replicated)
//   sub[t', :]      = ctx[t' * skip, :]                              L' = ceil(L / skip)   (skip <= 1: identity)
//   rows L' .. L'' - 1 are zero, L'' = L' rounded up to a multiple of n_downsample
//   x[n, t, :]      = sub_n[t, :] for t < L'_n, else 0;  input_sizes[n] = float(L''_n / T_max)
// Pure copies: bit-exact with the reference's NumPy / torch code.
#include "common.cuh"
#include "ctcb200.h"

namespace ctcb200 {
namespace {

__device__ __forceinline__ int out_length(long long L, int skip, int n_down, int* valid_rows) {
    long long ls = (skip <= 1) ? L : (L + skip - 1) / skip;
    *valid_rows = static_cast<int>(ls);
    if (n_down > 1 && ls % n_down != 0) ls += n_down - ls % n_down;
    return static_cast<int>(ls);
}

__global__ void __launch_bounds__(256)
assemble_features_kernel(const float* __restrict__ feat, const long long* __restrict__ offsets, int N, int F, int left, int right,
                         int skip, int n_down, int T_max, float* __restrict__ x, float* __restrict__ input_sizes) {
    const int Fo = F * (left + right + 1);
    const long long total = static_cast<long long>(N) * T_max * Fo;
    for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
         e += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int col = static_cast<int>(e % Fo);
        const long long row = e / Fo;
        const int t = static_cast<int>(row % T_max), n = static_cast<int>(row / T_max);
        const long long o0 = offsets[n], L = offsets[n + 1] - o0;
        int valid;
        const int Lpad = out_length(L, skip, n_down, &valid);
        float v = 0.0f;
        if (t < valid) {
            const int c = col / F, f = col - c * F;
            long long src_t = static_cast<long long>(t) * (skip <= 1 ? 1 : skip) + (c - left);
            src_t = src_t < 0 ? 0 : (src_t > L - 1 ? L - 1 : src_t);
            v = feat[(o0 + src_t) * F + f];
        }
        x[e] = v;
        if (col == 0 && t == 0) input_sizes[n] = static_cast<float>(static_cast<double>(Lpad) / static_cast<double>(T_max));
    }
}

__global__ void pad_labels_kernel(const long long* __restrict__ labels, const long long* __restrict__ offsets, int N, int S_max,
                                  long long* __restrict__ targets, long long* __restrict__ target_sizes) {
    const long long total = static_cast<long long>(N) * S_max;
    for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
         e += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int s = static_cast<int>(e % S_max), n = static_cast<int>(e / S_max);
        const long long o0 = offsets[n], S = offsets[n + 1] - o0;
        targets[e] = (s < S) ? labels[o0 + s] : 0;
        if (s == 0) target_sizes[n] = S;
    }
}

}  // namespace
}  // namespace ctcb200

using namespace ctcb200;

extern "C" CTCB200_API int ctcb200_assemble_features(const float* feat, const int64_t* offsets, int N, int F, int left, int right,
                                                     int skip, int n_downsample, int T_max, float* x, float* input_sizes,
                                                     ctcb200_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    CTCB_REQUIRE(N > 0 && F > 0 && T_max > 0, "assemble_features: empty N=%d F=%d T_max=%d", N, F, T_max);
    CTCB_REQUIRE(left >= 0 && right >= 0 && skip >= 0 && n_downsample >= 1, "assemble_features: bad context (%d, %d) skip %d downsample %d",
                 left, right, skip, n_downsample);
    const long long total = static_cast<long long>(N) * T_max * F * (left + right + 1);
    long long blocks = (total + 1023) / 1024;
    const long long cap = static_cast<long long>(device_sm_count()) * 8;
    if (blocks > cap) blocks = cap;
    assemble_features_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(feat, reinterpret_cast<const long long*>(offsets), N, F, left,
                                                                          right, skip, n_downsample, T_max, x, input_sizes);
    CTCB_LAUNCH_CHECK();
    return OK;
}

extern "C" CTCB200_API int ctcb200_pad_labels(const int64_t* labels, const int64_t* offsets, int N, int S_max, int64_t* targets,
                                              int64_t* target_sizes, ctcb200_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    CTCB_REQUIRE(N > 0 && S_max > 0, "pad_labels: empty N=%d S_max=%d", N, S_max);
    const long long total = static_cast<long long>(N) * S_max;
    int blocks = static_cast<int>((total + 255) / 256);
    if (blocks > 1024) blocks = 1024;
    pad_labels_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<const long long*>(labels), reinterpret_cast<const long long*>(offsets), N,
                                                S_max, reinterpret_cast<long long*>(targets), reinterpret_cast<long long*>(target_sizes));
    CTCB_LAUNCH_CHECK();
    return OK;
}
