// K9 — the optional 2-D convolution front of the acoustic model (LayerCNN, timit/models/model_ctc.py:38-68,
// applied at model_ctc.py:148): Conv2d(bias) -> BatchNorm2d -> ReLU [-> Dropout].
//
// The convolutions (3x3, 1 or 32 input channels, 32 output channels in the shipped config) are direct fp32 kernels working
// from shared-memory tiles (forward, weight gradient, data gradient; round 1 lowered them to im2col + the 128-wide tensor-core
// GEMM tiles, mostly padding: 3.7 ms per cfg3 step). Outputs are rows y[M = N*Ho*Wo, Cout]; BatchNorm2d statistics are
// per-channel over the M rows, i.e. exactly the row-statistics kernels of elementwise.cu; ReLU is fused with the BatchNorm
// apply. Activations are kept channel-last ([N,H,W,C] fp32) between blocks; the last block writes [N,H,C,W] so that the RNN
// stack sees the reference's feature order c*F' + f (model_ctc.py:153-158).
#include "common.cuh"
#include "ctcb200.h"

namespace ctcb200 {
namespace {

struct ConvGeom {
    int N, Hi, Wi, Cin, Ho, Wo, kh, kw, sh, sw, ph, pw;
};

// ---- direct fp32 convolution (forward, weight gradient, data gradient) ------------------------------------------------
// The conv front is 0.5 % of the model's FLOPs and its tensors are tall and thin (Cout = 32, K = 9 or 288), so it runs on the
// fp32 CUDA cores from shared-memory tiles instead of being padded into 128-wide tensor-core tiles: one CTA stages the input
// rows a band of TH output rows needs (zero-padded halo, channel-last, coalesced row copies) plus the whole weight tensor
// (<= 36 KB), and computes from shared memory. Exact fp32 arithmetic like the reference's nn.Conv2d (model_ctc.py:46-50).
constexpr int CONV_THREADS = 256;

struct Band {
    int TH;        // output rows per band
    int IH;        // input rows staged per band: (TH - 1) * sh + kh
    int IW;        // padded input row width in pixels: Wi + 2 * pw
    int bands;     // bands per image
};
__host__ __device__ inline Band make_band(const ConvGeom& g, int TH) {
    Band b;
    b.TH = TH;
    b.IH = (TH - 1) * g.sh + g.kh;
    b.IW = g.Wi + 2 * g.pw;
    b.bands = (g.Ho + TH - 1) / TH;
    return b;
}

// input rows [hi0, hi0 + IH) of image n -> xs[IH][IW][Cin], zero outside the image
__device__ __forceinline__ void stage_input_band(const float* __restrict__ x, float* __restrict__ xs, const ConvGeom& g,
                                                 const Band& b, int n, int hi0) {
    const int row_elems = b.IW * g.Cin, in_row = g.Wi * g.Cin, pad = g.pw * g.Cin;
    for (int e = threadIdx.x; e < b.IH * row_elems; e += CONV_THREADS) {
        const int rr = e / row_elems, ce = e - rr * row_elems;
        const int hi = hi0 + rr, src = ce - pad;
        float v = 0.0f;
        if (hi >= 0 && hi < g.Hi && src >= 0 && src < in_row)
            v = __ldg(x + (static_cast<long long>(n) * g.Hi + hi) * in_row + src);
        xs[e] = v;
    }
}

// y[m, o] = bias[o] + sum_{r,s,c} x[n, ho*sh-ph+r, wo*sw-pw+s, c] * w[o, c, r, s];  m = (n, ho, wo)
template <int CG>
__global__ void __launch_bounds__(CONV_THREADS)
conv2d_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                  float* __restrict__ y, ConvGeom g, int Cout, Band b) {
    extern __shared__ float smem_f[];
    const int K = g.kh * g.kw * g.Cin;
    float* ws = smem_f;                 // [K][Cout], k = (r, s, c)
    float* xs = ws + K * Cout;          // [IH][IW][Cin]
    for (int e = threadIdx.x; e < K * Cout; e += CONV_THREADS) {
        const int k = e / Cout, o = e - k * Cout;
        const int c = k % g.Cin, s = (k / g.Cin) % g.kw, r = k / (g.Cin * g.kw);
        ws[e] = __ldg(w + ((static_cast<long long>(o) * g.Cin + c) * g.kh + r) * g.kw + s);
    }
    const int groups = Cout / CG;
    for (int tile = blockIdx.x; tile < g.N * b.bands; tile += gridDim.x) {
        const int n = tile / b.bands, ho0 = (tile % b.bands) * b.TH;
        __syncthreads();
        stage_input_band(x, xs, g, b, n, ho0 * g.sh - g.ph);
        __syncthreads();
        const int rows = min(b.TH, g.Ho - ho0);
        for (int it = threadIdx.x; it < rows * g.Wo * groups; it += CONV_THREADS) {
            const int og = it % groups, p = it / groups;
            const int wo = p % g.Wo, hl = p / g.Wo;
            float acc[CG];
#pragma unroll
            for (int q = 0; q < CG; ++q) acc[q] = bias ? __ldg(bias + og * CG + q) : 0.0f;
            for (int r = 0; r < g.kh; ++r) {
                const float* xrow = xs + ((hl * g.sh + r) * b.IW + wo * g.sw) * g.Cin;   // pixel (wo*sw - pw + s) sits at padded index wo*sw + s
                const float* wrow = ws + (r * g.kw * g.Cin) * Cout + og * CG;
                for (int sc = 0; sc < g.kw * g.Cin; ++sc) {
                    const float xv = xrow[sc];
#pragma unroll
                    for (int q = 0; q < CG; ++q) acc[q] = fmaf(xv, wrow[sc * Cout + q], acc[q]);
                }
            }
            float* dst = y + ((static_cast<long long>(n) * g.Ho + ho0 + hl) * g.Wo + wo) * Cout + og * CG;
#pragma unroll
            for (int q = 0; q < CG; ++q) dst[q] = acc[q];
        }
    }
}

// partial[block][o][c][r][s] = sum over the block's bands of dy[m, o] * x[n, ho*sh-ph+r, wo*sw-pw+s, c]
__global__ void __launch_bounds__(CONV_THREADS)
conv2d_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ partial, ConvGeom g,
                    int Cout, Band b) {
    extern __shared__ float smem_f[];
    const int K = g.kh * g.kw * g.Cin;
    float* xs = smem_f;                                  // [IH][IW][Cin]
    float* dys = xs + b.IH * b.IW * g.Cin;               // [TH * Wo][Cout]
    constexpr int MAX_OWN = 40;                          // outputs (o, k) owned by one thread: K * Cout / 256 <= 40
    float acc[MAX_OWN];
#pragma unroll
    for (int i = 0; i < MAX_OWN; ++i) acc[i] = 0.0f;
    const int total = K * Cout;
    for (int tile = blockIdx.x; tile < g.N * b.bands; tile += gridDim.x) {
        const int n = tile / b.bands, ho0 = (tile % b.bands) * b.TH;
        const int rows = min(b.TH, g.Ho - ho0);
        __syncthreads();
        stage_input_band(x, xs, g, b, n, ho0 * g.sh - g.ph);
        const float* dsrc = dy + (static_cast<long long>(n) * g.Ho + ho0) * g.Wo * Cout;
        for (int e = threadIdx.x; e < rows * g.Wo * Cout; e += CONV_THREADS) dys[e] = __ldg(dsrc + e);
        __syncthreads();
#pragma unroll
        for (int i = 0; i < MAX_OWN; ++i) {
            const int idx = threadIdx.x + i * CONV_THREADS;   // (k, o) with o fastest: a warp shares k, lanes differ in o
            if (idx < total) {
                const int k = idx / Cout, o = idx - k * Cout;
                const int c = k % g.Cin, s = (k / g.Cin) % g.kw, r = k / (g.Cin * g.kw);
                float a = acc[i];
                for (int hl = 0; hl < rows; ++hl) {
                    const float* xrow = xs + ((hl * g.sh + r) * b.IW + s) * g.Cin + c;
                    const float* drow = dys + hl * g.Wo * Cout + o;
                    for (int wo = 0; wo < g.Wo; ++wo) a = fmaf(drow[wo * Cout], xrow[wo * g.sw * g.Cin], a);
                }
                acc[i] = a;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < MAX_OWN; ++i) {
        const int idx = threadIdx.x + i * CONV_THREADS;
        if (idx < total) {
            const int k = idx / Cout, o = idx - k * Cout;
            const int c = k % g.Cin, s = (k / g.Cin) % g.kw, r = k / (g.Cin * g.kw);
            partial[static_cast<long long>(blockIdx.x) * total + ((static_cast<long long>(o) * g.Cin + c) * g.kh + r) * g.kw + s] = acc[i];
        }
    }
}

// dw[e] = sum over blocks of partial[block][e] (fixed order: deterministic)
__global__ void conv2d_wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, int blocks, int total) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    float a = 0.0f;
    for (int bk = 0; bk < blocks; ++bk) a += partial[static_cast<long long>(bk) * total + e];
    dw[e] = a;
}

// dx[n, hi, wi, c] = sum over (r, s) with ho = (hi + ph - r) / sh, wo = (wi + pw - s) / sw integral and in range, and o:
//                    dy[(n, ho, wo), o] * w[o, c, r, s]
template <int CG>
__global__ void __launch_bounds__(CONV_THREADS)
conv2d_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w, float* __restrict__ dx, ConvGeom g, int Cout,
                    int TH_in) {
    extern __shared__ float smem_f[];
    const int taps = g.kh * g.kw;
    float* ws = smem_f;                           // [tap][o][c]
    float* dys = ws + taps * Cout * g.Cin;        // [rows of ho][Wo][Cout]
    for (int e = threadIdx.x; e < taps * Cout * g.Cin; e += CONV_THREADS) {
        const int c = e % g.Cin, o = (e / g.Cin) % Cout, tap = e / (g.Cin * Cout);
        const int r = tap / g.kw, s = tap % g.kw;
        ws[e] = __ldg(w + ((static_cast<long long>(o) * g.Cin + c) * g.kh + r) * g.kw + s);
    }
    const int bands = (g.Hi + TH_in - 1) / TH_in;
    const int groups = g.Cin / CG;
    for (int tile = blockIdx.x; tile < g.N * bands; tile += gridDim.x) {
        const int n = tile / bands, hi0 = (tile % bands) * TH_in;
        const int rows_in = min(TH_in, g.Hi - hi0);
        // output rows that can touch input rows [hi0, hi0 + rows_in): ho*sh - ph + r = hi
        int ho_lo = (hi0 + g.ph - (g.kh - 1) + g.sh - 1) / g.sh;
        if (hi0 + g.ph - (g.kh - 1) < 0) ho_lo = 0;
        int ho_hi = (hi0 + rows_in - 1 + g.ph) / g.sh;
        if (ho_hi > g.Ho - 1) ho_hi = g.Ho - 1;
        const int nrows = ho_hi - ho_lo + 1;
        __syncthreads();
        if (nrows > 0) {
            const float* dsrc = dy + (static_cast<long long>(n) * g.Ho + ho_lo) * g.Wo * Cout;
            for (int e = threadIdx.x; e < nrows * g.Wo * Cout; e += CONV_THREADS) dys[e] = __ldg(dsrc + e);
        }
        __syncthreads();
        for (int it = threadIdx.x; it < rows_in * g.Wi * groups; it += CONV_THREADS) {
            const int cg = it % groups, p = it / groups;
            const int wi = p % g.Wi, hl = p / g.Wi;
            const int hi = hi0 + hl;
            float acc[CG];
#pragma unroll
            for (int q = 0; q < CG; ++q) acc[q] = 0.0f;
            for (int r = 0; r < g.kh; ++r) {
                const int hn = hi + g.ph - r;
                if (hn < 0 || hn % g.sh != 0) continue;
                const int ho = hn / g.sh;
                if (ho < ho_lo || ho > ho_hi) continue;
                for (int s = 0; s < g.kw; ++s) {
                    const int wn = wi + g.pw - s;
                    if (wn < 0 || wn % g.sw != 0) continue;
                    const int wo = wn / g.sw;
                    if (wo >= g.Wo) continue;
                    const float* drow = dys + ((ho - ho_lo) * g.Wo + wo) * Cout;
                    const float* wtap = ws + (r * g.kw + s) * Cout * g.Cin + cg * CG;
                    for (int o = 0; o < Cout; ++o) {
                        const float dv = drow[o];
#pragma unroll
                        for (int q = 0; q < CG; ++q) acc[q] = fmaf(dv, wtap[o * g.Cin + q], acc[q]);
                    }
                }
            }
            float* dst = dx + ((static_cast<long long>(n) * g.Hi + hi) * g.Wi + wi) * g.Cin + cg * CG;
#pragma unroll
            for (int q = 0; q < CG; ++q) dst[q] = acc[q];
        }
    }
}

constexpr int CONV_TH = 8;        // output rows per band (forward / weight gradient)
constexpr int CONV_TH_IN = 16;    // input rows per band (data gradient)

int conv_grid(long long tiles, int per_sm) {
    long long cap = static_cast<long long>(device_sm_count()) * per_sm;
    return static_cast<int>(tiles < cap ? (tiles < 1 ? 1 : tiles) : cap);
}

// y[m, c] += bias[c]
__global__ void add_bias_rows_kernel(float* __restrict__ y, const float* __restrict__ bias, long long total, int C) {
    for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
         e += static_cast<long long>(gridDim.x) * blockDim.x)
        y[e] += bias[e % C];
}

// activation selector shared by the forward / backward kernels: the reference builds LayerCNN with
// supported_activate = {relu, tanh, sigmoid} (train_ctc.py:21, model_ctc.py:50); 3 = identity (layout change only)
enum { ACT_RELU = 0, ACT_TANH = 1, ACT_SIGMOID = 2, ACT_IDENTITY = 3 };
__device__ __forceinline__ float act_fwd(float v, int act) {
    if (act == ACT_RELU) return fmaxf(v, 0.0f);
    if (act == ACT_TANH) return tanhf(v);
    if (act == ACT_SIGMOID) return 1.0f / (1.0f + expf(-v));
    return v;
}
// derivative expressed through the activation's OUTPUT a (what the forward pass keeps)
__device__ __forceinline__ float act_grad(float a, int act) {
    if (act == ACT_RELU) return a > 0.0f ? 1.0f : 0.0f;
    if (act == ACT_TANH) return 1.0f - a * a;
    if (act == ACT_SIGMOID) return a * (1.0f - a);
    return 1.0f;
}

// a(n, h, w, c) = act(y[m, c] * scale[c] + shift[c]), written with explicit output strides
__global__ void __launch_bounds__(256)
affine_act_kernel(const float* __restrict__ y, const float* __restrict__ scale, const float* __restrict__ shift,
                  float* __restrict__ a, long long sn, long long sh, long long sw, long long sc, int N, int Ho, int Wo,
                  int C, int act) {
    const long long total = static_cast<long long>(N) * Ho * Wo * C;
    for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
         e += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(e % C);
        const long long m = e / C;
        const int w = static_cast<int>(m % Wo), h = static_cast<int>((m / Wo) % Ho);
        const long long n = m / (static_cast<long long>(Wo) * Ho);
        float v = y[e];
        if (scale) v = v * scale[c] + shift[c];
        a[n * sn + h * sh + w * sw + c * sc] = act_fwd(v, act);
    }
}

// dz[m, c] = act'(a(n,h,w,c)) * da(n,h,w,c)   (da and a share the strided layout)
__global__ void __launch_bounds__(256)
act_bwd_gather_kernel(const float* __restrict__ da, const float* __restrict__ a, float* __restrict__ dz, long long sn,
                      long long sh, long long sw, long long sc, int N, int Ho, int Wo, int C, int act) {
    const long long total = static_cast<long long>(N) * Ho * Wo * C;
    for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
         e += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(e % C);
        const long long m = e / C;
        const int w = static_cast<int>(m % Wo), h = static_cast<int>((m / Wo) % Ho);
        const long long n = m / (static_cast<long long>(Wo) * Ho);
        const long long o = n * sn + h * sh + w * sw + c * sc;
        dz[e] = act_grad(a ? a[o] : 1.0f, act) * da[o];
    }
}

// nn.MaxPool2d(pool) on channel-last activations (kernel = stride = pool, no padding, floor mode: model_ctc.py:53-54);
// idx keeps the position of the maximum inside its window (first maximum wins, like torch) for the backward pass
__global__ void __launch_bounds__(256)
maxpool2d_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, uint8_t* __restrict__ idx, int N, int H, int W, int C,
                     int kh, int kw, int Ho, int Wo) {
    const long long total = static_cast<long long>(N) * Ho * Wo * C;
    for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
         e += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(e % C);
        const long long m = e / C;
        const int wo = static_cast<int>(m % Wo), ho = static_cast<int>((m / Wo) % Ho);
        const long long n = m / (static_cast<long long>(Wo) * Ho);
        float best = -INFINITY;
        int arg = 0;
        for (int r = 0; r < kh; ++r)
            for (int q = 0; q < kw; ++q) {
                const float v = x[((n * H + ho * kh + r) * W + wo * kw + q) * C + c];
                if (v > best || (v != v && best == best)) { best = v; arg = r * kw + q; }
            }
        y[e] = best;
        idx[e] = static_cast<uint8_t>(arg);
    }
}
__global__ void __launch_bounds__(256)
maxpool2d_bwd_kernel(const float* __restrict__ dy, const uint8_t* __restrict__ idx, float* __restrict__ dx, int N, int H, int W,
                     int C, int kh, int kw, int Ho, int Wo) {
    const long long total = static_cast<long long>(N) * H * W * C;   // every input element written once (no atomics)
    for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
         e += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(e % C);
        const long long p = e / C;
        const int w = static_cast<int>(p % W), h = static_cast<int>((p / W) % H);
        const long long n = p / (static_cast<long long>(W) * H);
        const int ho = h / kh, wo = w / kw;
        float v = 0.0f;
        if (ho < Ho && wo < Wo) {
            const long long o = ((n * Ho + ho) * Wo + wo) * C + c;
            if (idx[o] == (h - ho * kh) * kw + (w - wo * kw)) v = dy[o];
        }
        dx[e] = v;
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

extern "C" CTCB200_API int ctcb200_conv2d_fwd(const float* x_nhwc, const float* w, const float* bias, float* y, int N, int Hi,
                                              int Wi, int Cin, int Cout, int Ho, int Wo, int kh, int kw, int sh, int sw, int ph,
                                              int pw, ctcb200_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    CTCB_REQUIRE(N > 0 && Ho > 0 && Wo > 0 && Cin > 0 && Cout > 0, "conv2d_fwd: empty geometry");
    ConvGeom g{N, Hi, Wi, Cin, Ho, Wo, kh, kw, sh, sw, ph, pw};
    const Band b = make_band(g, CONV_TH);
    const size_t smem = (static_cast<size_t>(kh) * kw * Cin * Cout + static_cast<size_t>(b.IH) * b.IW * Cin) * sizeof(float);
    CTCB_REQUIRE(smem <= 200 * 1024, "conv2d_fwd: tile needs %zu bytes of shared memory (Cin=%d Cout=%d Wi=%d)", smem, Cin, Cout, Wi);
    const int grid = conv_grid(static_cast<long long>(N) * b.bands, 4);
    if (Cout % 8 == 0) {
        CTCB_CUDA(cudaFuncSetAttribute(conv2d_fwd_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        conv2d_fwd_kernel<8><<<grid, CONV_THREADS, smem, stream>>>(x_nhwc, w, bias, y, g, Cout, b);
    } else {
        CTCB_CUDA(cudaFuncSetAttribute(conv2d_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        conv2d_fwd_kernel<1><<<grid, CONV_THREADS, smem, stream>>>(x_nhwc, w, bias, y, g, Cout, b);
    }
    CTCB_LAUNCH_CHECK();
    return OK;
}

extern "C" CTCB200_API int64_t ctcb200_conv2d_wgrad_ws_bytes(int Cin, int Cout, int kh, int kw) {
    return static_cast<int64_t>(device_sm_count()) * 2 * kh * kw * Cin * Cout * static_cast<int64_t>(sizeof(float));
}

extern "C" CTCB200_API int ctcb200_conv2d_wgrad(const float* x_nhwc, const float* dy, float* dw, void* ws, int N, int Hi, int Wi,
                                                int Cin, int Cout, int Ho, int Wo, int kh, int kw, int sh, int sw, int ph, int pw,
                                                ctcb200_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    ConvGeom g{N, Hi, Wi, Cin, Ho, Wo, kh, kw, sh, sw, ph, pw};
    const Band b = make_band(g, CONV_TH);
    const int total = kh * kw * Cin * Cout;
    CTCB_REQUIRE(total <= 40 * CONV_THREADS, "conv2d_wgrad: %d weights exceed the per-thread accumulator budget", total);
    const size_t smem = (static_cast<size_t>(b.IH) * b.IW * Cin + static_cast<size_t>(CONV_TH) * Wo * Cout) * sizeof(float);
    CTCB_REQUIRE(smem <= 200 * 1024, "conv2d_wgrad: tile needs %zu bytes of shared memory", smem);
    const int grid = conv_grid(static_cast<long long>(N) * b.bands, 2);
    CTCB_CUDA(cudaFuncSetAttribute(conv2d_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    conv2d_wgrad_kernel<<<grid, CONV_THREADS, smem, stream>>>(x_nhwc, dy, static_cast<float*>(ws), g, Cout, b);
    CTCB_LAUNCH_CHECK();
    conv2d_wgrad_reduce_kernel<<<(total + 255) / 256, 256, 0, stream>>>(static_cast<const float*>(ws), dw, grid, total);
    CTCB_LAUNCH_CHECK();
    return OK;
}

extern "C" CTCB200_API int ctcb200_conv2d_dgrad(const float* dy, const float* w, float* dx_nhwc, int N, int Hi, int Wi, int Cin,
                                                int Cout, int Ho, int Wo, int kh, int kw, int sh, int sw, int ph, int pw,
                                                ctcb200_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    ConvGeom g{N, Hi, Wi, Cin, Ho, Wo, kh, kw, sh, sw, ph, pw};
    const int max_rows = (CONV_TH_IN - 1 + kh - 1) / sh + 2;
    const size_t smem = (static_cast<size_t>(kh) * kw * Cout * Cin + static_cast<size_t>(max_rows) * Wo * Cout) * sizeof(float);
    CTCB_REQUIRE(smem <= 200 * 1024, "conv2d_dgrad: tile needs %zu bytes of shared memory", smem);
    const int bands = (Hi + CONV_TH_IN - 1) / CONV_TH_IN;
    const int grid = conv_grid(static_cast<long long>(N) * bands, 4);
    if (Cin % 8 == 0) {
        CTCB_CUDA(cudaFuncSetAttribute(conv2d_dgrad_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        conv2d_dgrad_kernel<8><<<grid, CONV_THREADS, smem, stream>>>(dy, w, dx_nhwc, g, Cout, CONV_TH_IN);
    } else {
        CTCB_CUDA(cudaFuncSetAttribute(conv2d_dgrad_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        conv2d_dgrad_kernel<1><<<grid, CONV_THREADS, smem, stream>>>(dy, w, dx_nhwc, g, Cout, CONV_TH_IN);
    }
    CTCB_LAUNCH_CHECK();
    return OK;
}

extern "C" CTCB200_API int ctcb200_add_bias_rows(float* y, const float* bias, int64_t R, int C, ctcb200_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    add_bias_rows_kernel<<<grid_for(R * C), 256, 0, stream>>>(y, bias, R * C, C);
    CTCB_LAUNCH_CHECK();
    return OK;
}

extern "C" CTCB200_API int ctcb200_affine_act(const float* y, const float* scale, const float* shift, float* a,
                                              int64_t sn, int64_t sh, int64_t sw, int64_t sc, int N, int Ho, int Wo,
                                              int C, int act, ctcb200_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    CTCB_REQUIRE(act >= 0 && act <= 3, "affine_act: activation code %d not in {0 relu, 1 tanh, 2 sigmoid, 3 identity}", act);
    affine_act_kernel<<<grid_for(static_cast<long long>(N) * Ho * Wo * C), 256, 0, stream>>>(y, scale, shift, a, sn, sh, sw, sc,
                                                                                           N, Ho, Wo, C, act);
    CTCB_LAUNCH_CHECK();
    return OK;
}

extern "C" CTCB200_API int ctcb200_act_bwd_gather(const float* da, const float* a, float* dz, int64_t sn, int64_t sh,
                                                  int64_t sw, int64_t sc, int N, int Ho, int Wo, int C, int act,
                                                  ctcb200_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    CTCB_REQUIRE(act >= 0 && act <= 3, "act_bwd_gather: activation code %d not in {0,1,2,3}", act);
    CTCB_REQUIRE(a != nullptr || act == 3, "act_bwd_gather: the activation output is needed for act=%d", act);
    act_bwd_gather_kernel<<<grid_for(static_cast<long long>(N) * Ho * Wo * C), 256, 0, stream>>>(da, a, dz, sn, sh, sw, sc, N,
                                                                                               Ho, Wo, C, act);
    CTCB_LAUNCH_CHECK();
    return OK;
}

extern "C" CTCB200_API int ctcb200_maxpool2d_fwd(const float* x_nhwc, float* y_nhwc, void* idx_u8, int N, int H, int W, int C,
                                                 int kh, int kw, ctcb200_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    CTCB_REQUIRE(kh > 0 && kw > 0 && kh * kw <= 256 && H >= kh && W >= kw, "maxpool2d_fwd: bad window %dx%d for %dx%d", kh, kw, H, W);
    const int Ho = H / kh, Wo = W / kw;
    maxpool2d_fwd_kernel<<<grid_for(static_cast<long long>(N) * Ho * Wo * C), 256, 0, stream>>>(
        x_nhwc, y_nhwc, static_cast<uint8_t*>(idx_u8), N, H, W, C, kh, kw, Ho, Wo);
    CTCB_LAUNCH_CHECK();
    return OK;
}

extern "C" CTCB200_API int ctcb200_maxpool2d_bwd(const float* dy_nhwc, const void* idx_u8, float* dx_nhwc, int N, int H, int W,
                                                 int C, int kh, int kw, ctcb200_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    CTCB_REQUIRE(kh > 0 && kw > 0 && H >= kh && W >= kw, "maxpool2d_bwd: bad window");
    const int Ho = H / kh, Wo = W / kw;
    maxpool2d_bwd_kernel<<<grid_for(static_cast<long long>(N) * H * W * C), 256, 0, stream>>>(
        dy_nhwc, static_cast<const uint8_t*>(idx_u8), dx_nhwc, N, H, W, C, kh, kw, Ho, Wo);
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
