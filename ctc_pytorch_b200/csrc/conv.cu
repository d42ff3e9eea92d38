// K9 — the optional 2-D convolution front of the acoustic model (LayerCNN, timit/models/model_ctc.py:38-68,
// applied at model_ctc.py:148): Conv2d(bias) -> BatchNorm2d -> ReLU [-> Dropout].
//
// The convolution is lowered to the tensor-core GEMM of gemm.cu: an im2col kernel writes the bf16 patch matrix
// cols[M = N*Ho*Wo, K = kh*kw*Cin] (and, for training, its transpose, the B operand of the weight-gradient
// GEMM); BatchNorm2d statistics are per-channel over the M rows, i.e. exactly the row-statistics kernels of
// elementwise.cu; ReLU is fused with the BatchNorm apply. Activations are kept channel-last ([N,H,W,C] fp32)
// between blocks so that GEMM outputs need no transposition; the last block writes [N,H,C,W] so that the RNN
// stack sees the reference's feature order c*F' + f (model_ctc.py:153-158).
// These kernels are HBM streaming kernels (the conv FLOPs are ~0.5 % of the model): coalesced rows, grid-stride.
#include "common.cuh"
#include "ctcb200.h"

namespace ctcb200 {
namespace {

struct ConvGeom {
    int N, Hi, Wi, Cin, Ho, Wo, kh, kw, sh, sw, ph, pw;
};

// cols[m, k] with m = (n, ho, wo), k = (r, s, c): x[n, ho*sh - ph + r, wo*sw - pw + s, c] (0 outside).
// transposed = 0 writes cols [M, Kp] (k fastest), transposed = 1 writes colsT [K, Mp] (m fastest).
__global__ void __launch_bounds__(256)
im2col_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, long long pitch, ConvGeom g, int transposed) {
    const long long M = static_cast<long long>(g.N) * g.Ho * g.Wo;
    const int K = g.kh * g.kw * g.Cin;
    const long long total = M * K;
    for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
         e += static_cast<long long>(gridDim.x) * blockDim.x) {
        long long m;
        int k;
        if (transposed) { k = static_cast<int>(e / M); m = e % M; }
        else { m = e / K; k = static_cast<int>(e % K); }
        const int wo = static_cast<int>(m % g.Wo);
        const int ho = static_cast<int>((m / g.Wo) % g.Ho);
        const int n = static_cast<int>(m / (static_cast<long long>(g.Wo) * g.Ho));
        const int c = k % g.Cin, s = (k / g.Cin) % g.kw, r = k / (g.Cin * g.kw);
        const int hi = ho * g.sh - g.ph + r, wi = wo * g.sw - g.pw + s;
        float v = 0.0f;
        if (hi >= 0 && hi < g.Hi && wi >= 0 && wi < g.Wi)
            v = x[((static_cast<long long>(n) * g.Hi + hi) * g.Wi + wi) * g.Cin + c];
        if (transposed) out[static_cast<long long>(k) * pitch + m] = __float2bfloat16(v);
        else out[m * pitch + k] = __float2bfloat16(v);
    }
}

// scatter-add of dcols [M, K] (fp32) back onto dx [N, Hi, Wi, Cin]
__global__ void __launch_bounds__(256)
col2im_kernel(const float* __restrict__ dcols, long long pitch, float* __restrict__ dx, ConvGeom g) {
    const long long M = static_cast<long long>(g.N) * g.Ho * g.Wo;
    const int K = g.kh * g.kw * g.Cin;
    const long long total = M * K;
    for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
         e += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long m = e / K;
        const int k = static_cast<int>(e % K);
        const int wo = static_cast<int>(m % g.Wo);
        const int ho = static_cast<int>((m / g.Wo) % g.Ho);
        const int n = static_cast<int>(m / (static_cast<long long>(g.Wo) * g.Ho));
        const int c = k % g.Cin, s = (k / g.Cin) % g.kw, r = k / (g.Cin * g.kw);
        const int hi = ho * g.sh - g.ph + r, wi = wo * g.sw - g.pw + s;
        if (hi >= 0 && hi < g.Hi && wi >= 0 && wi < g.Wi)
            atomicAdd(&dx[((static_cast<long long>(n) * g.Hi + hi) * g.Wi + wi) * g.Cin + c], dcols[m * pitch + k]);
    }
}

// torch weight [Cout, Cin, kh, kw] fp32 -> w_p bf16 [Cout, Kp] with k = (r, s, c) and w_pT bf16 [K, Coutp]
__global__ void pack_conv_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ w_p,
                                        __nv_bfloat16* __restrict__ w_pT, int Cout, int Cin, int kh, int kw, int Kp,
                                        int Coutp) {
    const int K = kh * kw * Cin;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < Cout * Kp; e += gridDim.x * blockDim.x) {
        const int o = e / Kp, k = e % Kp;
        float v = 0.0f;
        if (k < K) {
            const int c = k % Cin, s = (k / Cin) % kw, r = k / (Cin * kw);
            v = w[((static_cast<long long>(o) * Cin + c) * kh + r) * kw + s];
        }
        w_p[e] = __float2bfloat16(v);
        if (k < K && w_pT) w_pT[static_cast<long long>(k) * Coutp + o] = __float2bfloat16(v);
    }
}

// y[m, c] += bias[c]
__global__ void add_bias_rows_kernel(float* __restrict__ y, const float* __restrict__ bias, long long total, int C) {
    for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
         e += static_cast<long long>(gridDim.x) * blockDim.x)
        y[e] += bias[e % C];
}

// a(n, h, w, c) = relu(y[m, c] * scale[c] + shift[c]), written with explicit output strides
__global__ void __launch_bounds__(256)
affine_relu_kernel(const float* __restrict__ y, const float* __restrict__ scale, const float* __restrict__ shift,
                   float* __restrict__ a, long long sn, long long sh, long long sw, long long sc, int N, int Ho, int Wo,
                   int C) {
    const long long total = static_cast<long long>(N) * Ho * Wo * C;
    for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
         e += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(e % C);
        const long long m = e / C;
        const int w = static_cast<int>(m % Wo), h = static_cast<int>((m / Wo) % Ho);
        const long long n = m / (static_cast<long long>(Wo) * Ho);
        float v = y[e];
        if (scale) v = v * scale[c] + shift[c];
        a[n * sn + h * sh + w * sw + c * sc] = fmaxf(v, 0.0f);
    }
}

// dz[m, c] = a(n,h,w,c) > 0 ? da(n,h,w,c) : 0   (da and a share the strided layout)
__global__ void __launch_bounds__(256)
relu_bwd_gather_kernel(const float* __restrict__ da, const float* __restrict__ a, float* __restrict__ dz, long long sn,
                       long long sh, long long sw, long long sc, int N, int Ho, int Wo, int C) {
    const long long total = static_cast<long long>(N) * Ho * Wo * C;
    for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
         e += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(e % C);
        const long long m = e / C;
        const int w = static_cast<int>(m % Wo), h = static_cast<int>((m / Wo) % Ho);
        const long long n = m / (static_cast<long long>(Wo) * Ho);
        const long long o = n * sn + h * sh + w * sw + c * sc;
        dz[e] = a[o] > 0.0f ? da[o] : 0.0f;
    }
}

// out[c] = sum over rows of y[m, c]
__global__ void __launch_bounds__(256)
col_sum_kernel(const float* __restrict__ y, float* __restrict__ out, long long R, int C) {
    __shared__ float part[8][32];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx;
    float acc = 0.0f;
    if (c < C)
        for (long long r = static_cast<long long>(blockIdx.y) * 8 + ty; r < R; r += static_cast<long long>(gridDim.y) * 8)
            acc += y[r * C + c];
    part[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && c < C) {
        float s = 0.0f;
        for (int k = 0; k < 8; ++k) s += part[k][tx];
        atomicAdd(&out[c], s);
    }
}

int grid_for(long long work) {
    long long b = (work + 1023) / 1024;
    const long long cap = static_cast<long long>(device_sm_count()) * 8;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return static_cast<int>(b);
}

}  // namespace
}  // namespace ctcb200

using namespace ctcb200;

extern "C" CTCB200_API int ctcb200_conv_im2col(const float* x_nhwc, void* cols, int64_t pitch, int transposed, int N,
                                               int Hi, int Wi, int Cin, int Ho, int Wo, int kh, int kw, int sh, int sw,
                                               int ph, int pw, ctcb200_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    CTCB_REQUIRE(N > 0 && Ho > 0 && Wo > 0 && Cin > 0, "conv_im2col: empty geometry");
    ConvGeom g{N, Hi, Wi, Cin, Ho, Wo, kh, kw, sh, sw, ph, pw};
    const long long total = static_cast<long long>(N) * Ho * Wo * kh * kw * Cin;
    im2col_kernel<<<grid_for(total), 256, 0, stream>>>(x_nhwc, static_cast<__nv_bfloat16*>(cols), pitch, g, transposed);
    CTCB_LAUNCH_CHECK();
    return OK;
}

extern "C" CTCB200_API int ctcb200_conv_col2im(const float* dcols, int64_t pitch, float* dx_nhwc, int N, int Hi, int Wi,
                                               int Cin, int Ho, int Wo, int kh, int kw, int sh, int sw, int ph, int pw,
                                               ctcb200_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    ConvGeom g{N, Hi, Wi, Cin, Ho, Wo, kh, kw, sh, sw, ph, pw};
    CTCB_CUDA(cudaMemsetAsync(dx_nhwc, 0, sizeof(float) * static_cast<size_t>(N) * Hi * Wi * Cin, stream));
    const long long total = static_cast<long long>(N) * Ho * Wo * kh * kw * Cin;
    col2im_kernel<<<grid_for(total), 256, 0, stream>>>(dcols, pitch, dx_nhwc, g);
    CTCB_LAUNCH_CHECK();
    return OK;
}

extern "C" CTCB200_API int ctcb200_conv_pack_weight(const float* w, void* w_p, void* w_pT, int Cout, int Cin, int kh,
                                                    int kw, int Kp, int Coutp, ctcb200_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    CTCB_REQUIRE(Kp >= kh * kw * Cin && Coutp >= Cout, "conv_pack_weight: padded sizes too small");
    pack_conv_weight_kernel<<<grid_for(static_cast<long long>(Cout) * Kp), 256, 0, stream>>>(
        w, static_cast<__nv_bfloat16*>(w_p), static_cast<__nv_bfloat16*>(w_pT), Cout, Cin, kh, kw, Kp, Coutp);
    CTCB_LAUNCH_CHECK();
    return OK;
}

extern "C" CTCB200_API int ctcb200_add_bias_rows(float* y, const float* bias, int64_t R, int C, ctcb200_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    add_bias_rows_kernel<<<grid_for(R * C), 256, 0, stream>>>(y, bias, R * C, C);
    CTCB_LAUNCH_CHECK();
    return OK;
}

extern "C" CTCB200_API int ctcb200_affine_relu(const float* y, const float* scale, const float* shift, float* a,
                                               int64_t sn, int64_t sh, int64_t sw, int64_t sc, int N, int Ho, int Wo,
                                               int C, ctcb200_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    affine_relu_kernel<<<grid_for(static_cast<long long>(N) * Ho * Wo * C), 256, 0, stream>>>(y, scale, shift, a, sn, sh, sw,
                                                                                            sc, N, Ho, Wo, C);
    CTCB_LAUNCH_CHECK();
    return OK;
}

extern "C" CTCB200_API int ctcb200_relu_bwd_gather(const float* da, const float* a, float* dz, int64_t sn, int64_t sh,
                                                   int64_t sw, int64_t sc, int N, int Ho, int Wo, int C,
                                                   ctcb200_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    relu_bwd_gather_kernel<<<grid_for(static_cast<long long>(N) * Ho * Wo * C), 256, 0, stream>>>(da, a, dz, sn, sh, sw, sc, N,
                                                                                                Ho, Wo, C);
    CTCB_LAUNCH_CHECK();
    return OK;
}

extern "C" CTCB200_API int ctcb200_col_sum(const float* y, float* out, int64_t R, int C, ctcb200_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    CTCB_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * C, stream));
    long long rb = (R + 511) / 512;
    if (rb > 1024) rb = 1024;
    col_sum_kernel<<<dim3((C + 31) / 32, static_cast<unsigned>(rb)), 256, 0, stream>>>(y, out, R, C);
    CTCB_LAUNCH_CHECK();
    return OK;
}
