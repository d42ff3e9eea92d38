// K7 — CTC loss forward (log-alpha) and backward (log-beta + gradient), warp per utterance.
//
// Replaces the library call the reference makes at timit/steps/train_ctc.py:144,47,63
// (nn.CTCLoss(reduction='sum') forward + autograd backward). Semantics follow that call exactly:
// blank index configurable (reference uses 0), targets are a zero-padded 2-D int64 matrix
// (timit/utils/data_loader.py:125,140), frames t >= input_length get zero gradient, an infeasible
// alignment gives +inf loss (zero_infinity=False), and the gradient is emitted in the convention
// of torch's native kernel: d/dlog_probs = exp(lp) - exp(log-sum_{s:l'_s=c}(alpha+beta) + nll - lp).
//
// Mapping: the T-step dependency chains are the cost at N=32, so the alpha sweep (t = 0..T-1) and the beta sweep
// (t = T-1..0) of an utterance run CONCURRENTLY in two warps of one launch; each stores its (renormalised) history.
// Inside a warp the 2S+1 lattice states are blocked over the 32 lanes (KS consecutive states per lane) so the
// s-1 / s-2 neighbours are lane-local except at block edges, where one or two warp shuffles fetch them; each
// (t, n) row of log-probs is brought in with coalesced cp.async into a per-warp ring of 8 rows, 7 frames ahead.
// The gradient is then a fully parallel kernel (one warp per (t, n) row): posteriors exp(alpha+beta+nll-lp) are
// accumulated per class in shared memory, grad = exp(lp) - occupancy.
// Numerics: every RENORM frames a sweep subtracts its row maximum and accumulates the removed log-scale in double;
// the per-row scales are stored so the posterior exponent is formed as (a~ + b~ - lp) + float(A_t + B_t + nll).
#include "common.cuh"
#include "ctcb200.h"
#include <stdlib.h>

namespace ctcb200 {

namespace {

constexpr int WARPS_PER_BLOCK = 4;
constexpr int RENORM = 8;  // alpha / beta rows are renormalised every RENORM frames
constexpr int PF = 8;      // log-prob rows kept in flight per sweep (cp.async ring): hides the L2/HBM latency of a row
#define NEG_INF (-INFINITY)

__device__ __forceinline__ float lse2(float a, float b) {
    float m = fmaxf(a, b);
    if (m == NEG_INF) return NEG_INF;
    return m + logf(expf(a - m) + expf(b - m));
}
// Branch-free log-sum-exp of three terms on the dependent chain of the sweeps: ex2.approx / lg2.approx (one MUFU op
// each). The arguments of ex2 are <= 0 and only terms within ~20 of the maximum matter, where the scaled-argument
// rounding is < 1e-6 relative; the sum lies in [1, 3], where lg2.approx has ~2^-22 absolute error — both below the
// fp32 spacing of the renormalised alpha / beta values themselves. All-(-inf) inputs give -inf without a branch:
// ex2(-inf) = 0, lg2(0) = -inf.
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// The MUFU pipe (16 lanes per clock and SM quadrant) is what a sweep step costs at KS >= 2, so the maximum term is never
// exponentiated (it is exactly 1): three terms take 2 ex2 + 1 lg2, and the blank states — which have no skip transition, i.e.
// half of all states — take the two-term form with 1 ex2 + 1 lg2 (10 MUFU ops per lane and step at KS = 4 instead of 16).
__device__ __forceinline__ float lse3(float a, float b, float c) {
    constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
    const float hi = fmaxf(a, b), lo = fminf(a, b);
    const float m = fmaxf(hi, c), mid = fminf(hi, c);
    const float ms = (m == NEG_INF) ? 0.0f : m;
    const float sum = 1.0f + ex2_approx((mid - ms) * LOG2E) + ex2_approx((lo - ms) * LOG2E);
    return (m == NEG_INF) ? NEG_INF : fmaf(lg2_approx(sum), LN2, m);
}
__device__ __forceinline__ float lse2_fast(float a, float b) {
    constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
    const float m = fmaxf(a, b), lo = fminf(a, b);
    const float ms = (m == NEG_INF) ? 0.0f : m;
    const float sum = 1.0f + ex2_approx((lo - ms) * LOG2E);
    return (m == NEG_INF) ? NEG_INF : fmaf(lg2_approx(sum), LN2, m);
}

__device__ __forceinline__ void cp_async_4(float* smem_dst, const float* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int PENDING>
__device__ __forceinline__ void cp_async_wait_pending() { asm volatile("cp.async.wait_group %0;" ::"n"(PENDING) : "memory"); }

// Per-lane static description of the states this lane owns.
template <int KS>
struct LaneStates {
    int label[KS];     // l'_s
    bool valid[KS];    // s < L
    bool skip_in[KS];  // transition s-2 -> s allowed
    bool skip_out[KS]; // transition s -> s+2 allowed
};

template <int KS>
__device__ __forceinline__ void load_states(LaneStates<KS>& st, const int64_t* tgt, int S, int blank, int lane) {
    const int L = 2 * S + 1;
#pragma unroll
    for (int j = 0; j < KS; ++j) {
        int s = lane * KS + j;
        bool v = s < L;
        int lab = blank, lab_m2 = blank, lab_p2 = blank;
        if (v && (s & 1)) lab = static_cast<int>(tgt[(s - 1) >> 1]);
        if (v && (s & 1) && s >= 3) lab_m2 = static_cast<int>(tgt[(s - 3) >> 1]);
        if (v && (s & 1) && s + 2 < L) lab_p2 = static_cast<int>(tgt[(s + 1) >> 1]);
        st.label[j] = lab;
        st.valid[j] = v;
        st.skip_in[j] = v && (s & 1) && s >= 3 && lab != lab_m2;
        st.skip_out[j] = v && (s & 1) && (s + 2 < L) && lab != lab_p2;
    }
}

// ---------------------------------------------------------------------------------------------
// sweeps: warp 2k -> alpha of utterance k, warp 2k+1 -> beta of utterance k (same block)
// workspace per utterance: hist_a / hist_b [T][KS*32] floats (slot j*32+lane), off_a / off_b [T] doubles (log-scale
// removed from the stored row), nll_d (double)
// ---------------------------------------------------------------------------------------------
// The step loop is the whole cost at small N (one warp per scheduler: every instruction's latency is exposed, ncu: 177
// instructions and ~710 cycles per frame in the first version), so it is written for instruction count: frames are processed in
// aligned blocks of PF = RENORM = 8 with the body unrolled — ring slots, history offsets and the renormalisation point are then
// compile-time, pointers run instead of being recomputed, and the row stride of the ring is a template constant (RS floats; 0 =
// run-time C for C > 64); the frames before / after the aligned blocks take a generic loop with the same step code.
template <int KS, bool ALPHA_ONLY, int RS>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
ctc_sweep_kernel(const float* __restrict__ lp, const int64_t* __restrict__ targets, int64_t tgt_stride,
                 const int64_t* __restrict__ in_len, const int64_t* __restrict__ tgt_len, float* __restrict__ hist_a,
                 float* __restrict__ hist_b, double* __restrict__ off_a, double* __restrict__ off_b,
                 double* __restrict__ nll_d, float* __restrict__ nll, int T, int N, int C, int blank) {
    static_assert(PF == 8 && RENORM == 8, "the unrolled blocks assume ring depth = renormalisation period = 8");
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // ALPHA_ONLY (throughput form, see ctc_beta_grad_kernel): every warp is the alpha sweep of its own utterance
    const int n = ALPHA_ONLY ? blockIdx.x * WARPS_PER_BLOCK + warp : blockIdx.x * (WARPS_PER_BLOCK / 2) + (warp >> 1);
    const bool is_beta = ALPHA_ONLY ? false : (warp & 1);
    if (n >= N) return;
    const int RSTR = RS ? RS : C;            // ring row stride in floats
    float* rowbuf = smem + warp * PF * RSTR;  // ring of PF rows, row t in slot t & 7

    const int S = static_cast<int>(tgt_len[n]);
    const int L = 2 * S + 1;
    int Tn = static_cast<int>(in_len[n]);
    if (Tn > T) Tn = T;
    if (Tn <= 0) {
        if (lane == 0 && !is_beta) {
            nll[n] = (S == 0) ? 0.0f : INFINITY;
            nll_d[n] = (S == 0) ? 0.0 : static_cast<double>(INFINITY);
        }
        return;
    }
    LaneStates<KS> st;
    load_states<KS>(st, targets + n * tgt_stride, S, blank, lane);
    const long long row_stride = static_cast<long long>(N) * C;
    const float* lp_n = lp + static_cast<size_t>(n) * C;
    float* hist = (is_beta ? hist_b : hist_a) + static_cast<size_t>(n) * T * (KS * 32) + lane;   // + t*KS*32 + j*32
    double* offs = (is_beta ? off_b : off_a) + static_cast<size_t>(n) * T;
    double scale_acc = 0.0;  // log-scale removed so far (double: the stored rows stay O(100) however long the utterance)
    float a[KS];

    // one row of log-probs into a ring slot; exactly one commit group per call (empty when !pred) so that the group count
    // stays in step with the frame count
    const bool c0 = lane < C, c1 = lane + 32 < C;
    auto fetch_row = [&](float* dst, const float* src, bool pred) {
        if (pred) {
            if (c0) cp_async_4(dst + lane, src + lane);
            if (c1) cp_async_4(dst + lane + 32, src + lane + 32);
            if (RS == 0)
                for (int c = lane + 64; c < C; c += 32) cp_async_4(dst + c, src + c);
        }
        cp_async_commit();
    };
    auto renorm = [&]() {
        float m = NEG_INF;
#pragma unroll
        for (int j = 0; j < KS; ++j) m = fmaxf(m, a[j]);
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (m > NEG_INF) {
#pragma unroll
            for (int j = 0; j < KS; ++j) a[j] -= m;
            scale_acc += static_cast<double>(m);
        }
    };
    auto store_row = [&](float* h, double* o) {
        if (lane == 0) *o = scale_acc;
#pragma unroll
        for (int j = 0; j < KS; ++j) h[j * 32] = a[j];
    };

    if (!is_beta) {
        auto alpha_init = [&](const float* cur) {
#pragma unroll
            for (int j = 0; j < KS; ++j) a[j] = (st.valid[j] && lane * KS + j < 2) ? cur[st.label[j]] : NEG_INF;
        };
        auto alpha_step = [&](const float* cur) {
            float up1 = __shfl_up_sync(0xffffffffu, a[KS - 1], 1);
            float up2 = (KS >= 2) ? __shfl_up_sync(0xffffffffu, a[KS >= 2 ? KS - 2 : 0], 1)
                                  : __shfl_up_sync(0xffffffffu, a[0], 2);
            if (lane == 0) { up1 = NEG_INF; up2 = NEG_INF; }
            if (KS == 1 && lane == 1) up2 = NEG_INF;
            float nw[KS];
#pragma unroll
            for (int j = 0; j < KS; ++j) {
                float p1 = (j >= 1) ? a[j >= 1 ? j - 1 : 0] : up1;
                float p2 = (j >= 2) ? a[j >= 2 ? j - 2 : 0] : ((j == 1) ? up1 : up2);
                if (KS >= 2 && j == 0) p2 = up2;
                if (!st.skip_in[j]) p2 = NEG_INF;
                // even states are blanks (no skip transition): known at compile time when KS is even
                float v = (KS % 2 == 0 && (j & 1) == 0) ? lse2_fast(a[j], p1) : lse3(a[j], p1, p2);
                nw[j] = st.valid[j] ? (v + cur[st.label[j]]) : NEG_INF;
            }
#pragma unroll
            for (int j = 0; j < KS; ++j) a[j] = nw[j];
        };
        // rows 0 .. PF-2 in flight before the first frame; frame t tops the ring up with row t + PF - 1, whose slot held
        // row t - 1 (all lanes are done with it after the __syncwarp)
#pragma unroll
        for (int i = 0; i < PF - 1; ++i) fetch_row(rowbuf + i * RSTR, lp_n + i * row_stride, i < Tn);
        const float* next = lp_n + (PF - 1) * row_stride;   // row t + PF - 1 of the frame about to run
        const int full = Tn & ~(PF - 1);                    // frames in aligned blocks of 8
        for (int tb = 0; tb < full; tb += PF) {
#pragma unroll
            for (int i = 0; i < PF; ++i) {
                cp_async_wait_pending<PF - 2>();
                __syncwarp();
                fetch_row(rowbuf + ((i + PF - 1) % PF) * RSTR, next, tb + i + PF - 1 < Tn);
                next += row_stride;
                const float* cur = rowbuf + i * RSTR;
                if (i == 0 && tb == 0) alpha_init(cur);
                else alpha_step(cur);
                if (i == PF - 1) renorm();
                store_row(hist + i * (KS * 32), offs + i);
            }
            hist += PF * (KS * 32);
            offs += PF;
        }
        for (int t = full; t < Tn; ++t) {   // the last Tn % 8 frames
            cp_async_wait_pending<PF - 2>();
            __syncwarp();
            fetch_row(rowbuf + ((t + PF - 1) % PF) * RSTR, next, false);   // (nothing left to fetch: t + 7 >= Tn here)
            const float* cur = rowbuf + (t % PF) * RSTR;
            if (t == 0) alpha_init(cur);
            else alpha_step(cur);
            if (t == Tn - 1) renorm();
            store_row(hist, offs);
            hist += KS * 32;
            offs += 1;
        }
        // nll = -lse(alpha_{Tn-1}(L-1), alpha_{Tn-1}(L-2))
        float loc = NEG_INF;
#pragma unroll
        for (int j = 0; j < KS; ++j) {
            int s = lane * KS + j;
            if (s == L - 1 || s == L - 2) loc = lse2(loc, a[j]);
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) loc = lse2(loc, __shfl_xor_sync(0xffffffffu, loc, o));
        if (lane == 0) {
            const double v = -(static_cast<double>(loc) + scale_acc);
            nll[n] = static_cast<float>(v);
            nll_d[n] = v;
        }
    } else {
        auto beta_init = [&](const float* cur) {
#pragma unroll
            for (int j = 0; j < KS; ++j) {
                const int s = lane * KS + j;
                a[j] = (st.valid[j] && (s == L - 1 || s == L - 2)) ? cur[st.label[j]] : NEG_INF;
            }
        };
        auto beta_step = [&](const float* cur) {
            float dn1 = __shfl_down_sync(0xffffffffu, a[0], 1);
            float dn2 = (KS >= 2) ? __shfl_down_sync(0xffffffffu, a[KS >= 2 ? 1 : 0], 1)
                                  : __shfl_down_sync(0xffffffffu, a[0], 2);
            if (lane == 31) { dn1 = NEG_INF; dn2 = NEG_INF; }
            if (KS == 1 && lane == 30) dn2 = NEG_INF;
            float nw[KS];
#pragma unroll
            for (int j = 0; j < KS; ++j) {
                float n1 = (j + 1 < KS) ? a[j + 1 < KS ? j + 1 : 0] : dn1;
                float n2 = (j + 2 < KS) ? a[j + 2 < KS ? j + 2 : 0] : ((j + 1 < KS) ? dn1 : dn2);
                if (KS >= 2 && j == KS - 1) n2 = dn2;
                if (!st.skip_out[j]) n2 = NEG_INF;
                float v = (KS % 2 == 0 && (j & 1) == 0) ? lse2_fast(a[j], n1) : lse3(a[j], n1, n2);
                nw[j] = st.valid[j] ? (v + cur[st.label[j]]) : NEG_INF;
            }
#pragma unroll
            for (int j = 0; j < KS; ++j) a[j] = nw[j];
        };
        // rows Tn-1 .. Tn-PF+1 in flight before the first frame; frame t tops the ring up with row t - (PF - 1)
#pragma unroll
        for (int i = 0; i < PF - 1; ++i) {
            const int r = Tn - 1 - i;
            fetch_row(rowbuf + (r & (PF - 1)) * RSTR, lp_n + static_cast<long long>(r) * row_stride, r >= 0);
        }
        const float* next = lp_n + static_cast<long long>(Tn - PF) * row_stride;   // row t - (PF - 1) of the frame about to run
        const int full = Tn & ~(PF - 1);
        hist += static_cast<size_t>(Tn - 1) * (KS * 32);
        offs += Tn - 1;
        for (int t = Tn - 1; t >= full; --t) {   // the top Tn % 8 frames
            cp_async_wait_pending<PF - 2>();
            __syncwarp();
            fetch_row(rowbuf + ((t + 1) & (PF - 1)) * RSTR, next, t - (PF - 1) >= 0);
            next -= row_stride;
            const float* cur = rowbuf + (t & (PF - 1)) * RSTR;
            if (t == Tn - 1) beta_init(cur);
            else beta_step(cur);
            if ((t & (PF - 1)) == 0) renorm();   // (t == full: the lowest frame of the partial top block)
            store_row(hist, offs);
            hist -= KS * 32;
            offs -= 1;
        }
        // hist / offs now point at frame full - 1 = the top frame of the aligned blocks
        for (int tb = full - PF; tb >= 0; tb -= PF) {
            hist -= (PF - 1) * (KS * 32);   // -> frame tb
            offs -= PF - 1;
#pragma unroll
            for (int i = PF - 1; i >= 0; --i) {
                cp_async_wait_pending<PF - 2>();
                __syncwarp();
                fetch_row(rowbuf + ((i + 1) % PF) * RSTR, next, tb + i - (PF - 1) >= 0);
                next -= row_stride;
                const float* cur = rowbuf + i * RSTR;
                if (i == PF - 1 && tb + PF == Tn) beta_init(cur);
                else beta_step(cur);
                if (i == 0) renorm();
                store_row(hist + i * (KS * 32), offs + i);
            }
            hist -= KS * 32;                // -> frame tb - 1
            offs -= 1;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// gradient: one warp per (t, n) row, no dependency between rows
// ---------------------------------------------------------------------------------------------
template <int KS>
__global__ void __launch_bounds__(256)
ctc_grad_kernel(const float* __restrict__ lp, const int64_t* __restrict__ targets, int64_t tgt_stride,
                const int64_t* __restrict__ in_len, const int64_t* __restrict__ tgt_len, const float* __restrict__ hist_a,
                const float* __restrict__ hist_b, const double* __restrict__ off_a, const double* __restrict__ off_b,
                const double* __restrict__ nll_d, const float* __restrict__ grad_nll, float grad_scale,
                float* __restrict__ grad, int T, int N, int C, int blank) {
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long wpb = blockDim.x >> 5;
    float* occ = smem + warp * C;
    const long long rows = static_cast<long long>(T) * N;
    for (long long row = blockIdx.x * wpb + warp; row < rows; row += gridDim.x * wpb) {
        const int t = static_cast<int>(row / N), n = static_cast<int>(row % N);
        float* g_row = grad + row * C;
        int Tn = static_cast<int>(in_len[n]);
        if (Tn > T) Tn = T;
        if (t >= Tn) {  // frames past the utterance end: zero gradient (torch semantics)
            for (int c = lane; c < C; c += 32) g_row[c] = 0.0f;
            continue;
        }
        const int S = static_cast<int>(tgt_len[n]);
        LaneStates<KS> st;
        load_states<KS>(st, targets + n * tgt_stride, S, blank, lane);
        const float* lp_row = lp + row * C;
        for (int c = lane; c < C; c += 32) occ[c] = 0.0f;
        __syncwarp();
        const size_t h = (static_cast<size_t>(n) * T + t) * (KS * 32);
        const float shift = static_cast<float>(off_a[static_cast<size_t>(n) * T + t] + off_b[static_cast<size_t>(n) * T + t] +
                                               nll_d[n]);
        float blank_sum = 0.0f;
#pragma unroll
        for (int j = 0; j < KS; ++j) {
            if (st.valid[j]) {
                const float al = hist_a[h + j * 32 + lane], be = hist_b[h + j * 32 + lane];
                const float lpv = __ldg(lp_row + st.label[j]);
                const float g = expf((al + be - lpv) + shift);
                const int s = lane * KS + j;
                if (s & 1) atomicAdd(&occ[st.label[j]], g);
                else blank_sum += g;
            }
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) blank_sum += __shfl_xor_sync(0xffffffffu, blank_sum, o);
        if (lane == 0) atomicAdd(&occ[blank], blank_sum);
        __syncwarp();
        const float gscale = grad_scale * (grad_nll ? grad_nll[n] : 1.0f);
        for (int c = lane; c < C; c += 32) g_row[c] = (expf(__ldg(lp_row + c)) - occ[c]) * gscale;
        __syncwarp();
    }
}


// ---------------------------------------------------------------------------------------------
// Throughput form for large batches (N >= ctc_fused_min_batch): the alpha sweep alone (ctc_sweep_kernel<KS, true, RS>, one warp
// per utterance, history kept), then ONE kernel that runs the beta sweep and emits the gradient row of frame t the moment
// beta_t is known — no beta history, no separate gradient pass. HBM traffic per (t, n) row: log-probs twice (2 x 4C bytes),
// the alpha row once each way (2 x 128 KS bytes), the gradient once (4C bytes); the latency form above moves two histories
// each way and reads the log-probs three times. The chain of an utterance is twice as long (alpha, then beta), which is why
// small batches — where the T dependent steps are the cost — keep the concurrent sweeps.
// The arithmetic is the gradient kernel's, term for term: g_s = exp((alpha~ + beta~ - lp) + float(A_t + B_t + nll)).
// Ring slot r of a warp: [C_pad floats log-probs of row r | KS*32 floats alpha row r], both brought in with cp.async
// PF - 1 frames ahead of the sweep.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async_16(float* smem_dst, const float* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}

template <int KS, int RS>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
ctc_beta_grad_kernel(const float* __restrict__ lp, const int64_t* __restrict__ targets, int64_t tgt_stride,
                     const int64_t* __restrict__ in_len, const int64_t* __restrict__ tgt_len,
                     const float* __restrict__ hist_a, const double* __restrict__ off_a, const double* __restrict__ nll_d,
                     const float* __restrict__ grad_nll, float grad_scale, float* __restrict__ grad, int T, int N, int C,
                     int blank) {
    extern __shared__ __align__(16) float smem_bg[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = blockIdx.x * WARPS_PER_BLOCK + warp;
    if (n >= N) return;
    const int C_pad = RS ? RS : ((C + 3) & ~3);   // log-prob part of a ring slot (compile-time 64 floats for C <= 64)
    const int ROW = C_pad + KS * 32;
    float* ring = smem_bg + warp * (PF * ROW + C_pad);
    float* occ = ring + PF * ROW;

    const int S = static_cast<int>(tgt_len[n]);
    const int L = 2 * S + 1;
    int Tn = static_cast<int>(in_len[n]);
    if (Tn > T) Tn = T;
    const long long row_stride = static_cast<long long>(N) * C;
    const float* lp_n = lp + static_cast<size_t>(n) * C;
    float* g_n = grad + static_cast<size_t>(n) * C;
    // frames past the utterance end: zero gradient (torch semantics)
    for (int t = Tn > 0 ? Tn : 0; t < T; ++t)
        for (int c = lane; c < C; c += 32) g_n[t * row_stride + c] = 0.0f;
    if (Tn <= 0) return;
    LaneStates<KS> st;
    load_states<KS>(st, targets + n * tgt_stride, S, blank, lane);
    const float* hist = hist_a + static_cast<size_t>(n) * T * (KS * 32);
    const double* offs = off_a + static_cast<size_t>(n) * T;
    const double nllv = nll_d[n];
    const float gscale = grad_scale * (grad_nll ? grad_nll[n] : 1.0f);
    for (int c = lane; c < C; c += 32) occ[c] = 0.0f;
    const bool c0 = lane < C, c1 = lane + 32 < C;

    // log-prob row and alpha row of one frame into a ring slot; one commit group per call (empty when !pred)
    auto fetch_row = [&](float* dst, const float* src, const float* hsrc, bool pred) {
        if (pred) {
            if (c0) cp_async_4(dst + lane, src + lane);
            if (c1) cp_async_4(dst + lane + 32, src + lane + 32);
            if (RS == 0)
                for (int c = lane + 64; c < C; c += 32) cp_async_4(dst + c, src + c);
#pragma unroll
            for (int i = 0; i < (KS * 8 + 31) / 32; ++i)
                if (KS * 8 >= 32 || lane < KS * 8) cp_async_16(dst + C_pad + (i * 32 + lane) * 4, hsrc + (i * 32 + lane) * 4);
        }
        cp_async_commit();
    };

    double scale_acc = 0.0;
    float a[KS];
    auto beta_init = [&](const float* cur) {
#pragma unroll
        for (int j = 0; j < KS; ++j) {
            const int s = lane * KS + j;
            a[j] = (st.valid[j] && (s == L - 1 || s == L - 2)) ? cur[st.label[j]] : NEG_INF;
        }
    };
    auto beta_step = [&](const float* cur) {
        float dn1 = __shfl_down_sync(0xffffffffu, a[0], 1);
        float dn2 = (KS >= 2) ? __shfl_down_sync(0xffffffffu, a[KS >= 2 ? 1 : 0], 1)
                              : __shfl_down_sync(0xffffffffu, a[0], 2);
        if (lane == 31) { dn1 = NEG_INF; dn2 = NEG_INF; }
        if (KS == 1 && lane == 30) dn2 = NEG_INF;
        float nw[KS];
#pragma unroll
        for (int j = 0; j < KS; ++j) {
            float n1 = (j + 1 < KS) ? a[j + 1 < KS ? j + 1 : 0] : dn1;
            float n2 = (j + 2 < KS) ? a[j + 2 < KS ? j + 2 : 0] : ((j + 1 < KS) ? dn1 : dn2);
            if (KS >= 2 && j == KS - 1) n2 = dn2;
            if (!st.skip_out[j]) n2 = NEG_INF;
            float v = (KS % 2 == 0 && (j & 1) == 0) ? lse2_fast(a[j], n1) : lse3(a[j], n1, n2);
            nw[j] = st.valid[j] ? (v + cur[st.label[j]]) : NEG_INF;
        }
#pragma unroll
        for (int j = 0; j < KS; ++j) a[j] = nw[j];
    };
    auto renorm = [&]() {
        float m = NEG_INF;
#pragma unroll
        for (int j = 0; j < KS; ++j) m = fmaxf(m, a[j]);
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (m > NEG_INF) {
#pragma unroll
            for (int j = 0; j < KS; ++j) a[j] -= m;
            scale_acc += static_cast<double>(m);
        }
    };
    // state posteriors of a frame summed per class, then its gradient row
    auto emit = [&](const float* cur, double offa, float* g_row) {
        const float* al = cur + C_pad;
        const float shift = static_cast<float>(offa + scale_acc + nllv);
        float blank_sum = 0.0f;
#pragma unroll
        for (int j = 0; j < KS; ++j) {
            if (st.valid[j]) {
                const float g = expf((al[j * 32 + lane] + a[j] - cur[st.label[j]]) + shift);
                if ((lane * KS + j) & 1) atomicAdd(&occ[st.label[j]], g);
                else blank_sum += g;
            }
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) blank_sum += __shfl_xor_sync(0xffffffffu, blank_sum, o);
        if (lane == 0) atomicAdd(&occ[blank], blank_sum);
        __syncwarp();
        if (c0) { g_row[lane] = (expf(cur[lane]) - occ[lane]) * gscale; occ[lane] = 0.0f; }
        if (c1) { g_row[lane + 32] = (expf(cur[lane + 32]) - occ[lane + 32]) * gscale; occ[lane + 32] = 0.0f; }
        if (RS == 0)
            for (int c = lane + 64; c < C; c += 32) {
                g_row[c] = (expf(cur[c]) - occ[c]) * gscale;
                occ[c] = 0.0f;
            }
    };

    // rows Tn-1 .. Tn-PF+1 in flight before the first frame; frame t tops the ring up with row t - (PF - 1) (slot (t + 1) & 7,
    // whose row t + 1 every lane is done with after the __syncwarp)
#pragma unroll
    for (int i = 0; i < PF - 1; ++i) {
        const int r = Tn - 1 - i;
        fetch_row(ring + (r & (PF - 1)) * ROW, lp_n + static_cast<long long>(r) * row_stride,
                  hist + static_cast<long long>(r) * (KS * 32), r >= 0);
    }
    const float* next = lp_n + static_cast<long long>(Tn - PF) * row_stride;      // row t - (PF - 1) of the frame about to run
    const float* hnext = hist + static_cast<long long>(Tn - PF) * (KS * 32);
    float* g_row = g_n + static_cast<long long>(Tn - 1) * row_stride;
    const double* op = offs + (Tn - 1);
    double offa_next = *op;
    const int full = Tn & ~(PF - 1);
    for (int t = Tn - 1; t >= full; --t) {   // the top Tn % 8 frames
        cp_async_wait_pending<PF - 2>();
        __syncwarp();   // row t has landed for every lane; every lane is done with row t + 1 and with its occupancy sums
        fetch_row(ring + ((t + 1) & (PF - 1)) * ROW, next, hnext, t - (PF - 1) >= 0);
        next -= row_stride;
        hnext -= KS * 32;
        const double offa = offa_next;
        if (t > 0) offa_next = *--op;
        const float* cur = ring + (t & (PF - 1)) * ROW;
        if (t == Tn - 1) beta_init(cur);
        else beta_step(cur);
        if ((t & (PF - 1)) == 0) renorm();
        emit(cur, offa, g_row);
        g_row -= row_stride;
    }
    for (int tb = full - PF; tb >= 0; tb -= PF) {
#pragma unroll
        for (int i = PF - 1; i >= 0; --i) {
            cp_async_wait_pending<PF - 2>();
            __syncwarp();
            fetch_row(ring + ((i + 1) % PF) * ROW, next, hnext, tb + i - (PF - 1) >= 0);
            next -= row_stride;
            hnext -= KS * 32;
            const double offa = offa_next;
            if (i > 0 || tb > 0) offa_next = *--op;
            const float* cur = ring + i * ROW;
            if (i == PF - 1 && tb + PF == Tn) beta_init(cur);
            else beta_step(cur);
            if (i == 0) renorm();
            emit(cur, offa, g_row);
            g_row -= row_stride;
        }
    }
}

// Batch size from which the throughput form is used (ctcb200_ctc_set_fused_min_batch; CTCB200_CTC_FUSED_MIN_N at load time).
int& ctc_fused_min_batch() {
    static int v = [] {
        const char* e = getenv("CTCB200_CTC_FUSED_MIN_N");
        return e ? atoi(e) : 2048;
    }();
    return v;
}
bool ctc_fused(int N) { return N >= ctc_fused_min_batch(); }

int ks_for(int max_target_len) {
    int L = 2 * max_target_len + 1;
    int k = (L + 31) / 32;
    int ks = 1;
    while (ks < k) ks <<= 1;
    return ks;
}

// workspace layout (floats): hist_a | hist_b | then doubles: off_a [N*T] | off_b [N*T] | nll_d [N]
struct CtcWs {
    float* hist_a; float* hist_b; double* off_a; double* off_b; double* nll_d;
};
int64_t hist_floats(int T, int N, int ks) { return ((static_cast<int64_t>(N) * T * ks * 32 + 1) / 2) * 2; }
CtcWs carve(float* ws, int T, int N, int ks) {
    CtcWs w;
    w.hist_a = ws;
    w.hist_b = ws + hist_floats(T, N, ks);
    w.off_a = reinterpret_cast<double*>(ws + 2 * hist_floats(T, N, ks));
    w.off_b = w.off_a + static_cast<int64_t>(N) * T;
    w.nll_d = w.off_b + static_cast<int64_t>(N) * T;
    return w;
}

}  // namespace

}  // namespace ctcb200

using namespace ctcb200;

extern "C" CTCB200_API int64_t ctcb200_ctc_workspace_floats(int T, int N, int max_target_len) {
    int ks = ks_for(max_target_len);
    return 2 * hist_floats(T, N, ks) + 2 * (2 * static_cast<int64_t>(N) * T + N) + 2;
}

extern "C" CTCB200_API int ctcb200_ctc_set_fused_min_batch(int min_batch) {
    const int prev = ctc_fused_min_batch();
    if (min_batch >= 0) ctc_fused_min_batch() = min_batch;
    return prev;
}

extern "C" CTCB200_API int ctcb200_ctc_loss_fwd(const float* log_probs, const int64_t* targets, int64_t target_stride,
                                    const int64_t* input_lengths, const int64_t* target_lengths, int T, int N,
                                    int C, int max_target_len, int blank, float* alpha_ws, float* nll,
                                    ctcb200_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    CTCB_REQUIRE(T > 0 && N > 0 && C > 0, "ctc_loss_fwd: empty shape T=%d N=%d C=%d", T, N, C);
    CTCB_REQUIRE(blank >= 0 && blank < C, "ctc_loss_fwd: blank %d out of range [0,%d)", blank, C);
    CTCB_REQUIRE((reinterpret_cast<uintptr_t>(alpha_ws) & 15) == 0, "ctc_loss_fwd: workspace must be 16-byte aligned");
    int ks = ks_for(max_target_len);
    CTCB_REQUIRE(ks <= 16, "ctc_loss_fwd: target length %d exceeds the supported maximum 255", max_target_len);
    const bool fused = ctc_fused(N);   // throughput form: alpha sweeps only here, beta + gradient in ctcb200_ctc_loss_bwd
    const int utt_per_block = fused ? WARPS_PER_BLOCK : WARPS_PER_BLOCK / 2;
    dim3 grid((N + utt_per_block - 1) / utt_per_block), block(WARPS_PER_BLOCK * 32);
    const bool rs64 = C <= 64;   // ring row stride: compile-time 64 floats, or the run-time class count
    size_t smem = static_cast<size_t>(WARPS_PER_BLOCK) * PF * (rs64 ? 64 : C) * sizeof(float);
    CTCB_REQUIRE(smem <= 200 * 1024, "ctc_loss_fwd: class count %d too large for the row buffer", C);
    CtcWs w = carve(alpha_ws, T, N, ks);
#define LAUNCH_A3(KS, AO, RS)                                                                                           \
    CTCB_CUDA(cudaFuncSetAttribute(ctc_sweep_kernel<KS, AO, RS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    ctc_sweep_kernel<KS, AO, RS><<<grid, block, smem, stream>>>(log_probs, targets, target_stride, input_lengths,       \
                                                                target_lengths, w.hist_a, w.hist_b, w.off_a, w.off_b,   \
                                                                w.nll_d, nll, T, N, C, blank)
#define LAUNCH_A2(KS, AO)      \
    if (rs64) {                \
        LAUNCH_A3(KS, AO, 64); \
    } else {                   \
        LAUNCH_A3(KS, AO, 0);  \
    }
#define LAUNCH_A(KS)          \
    if (fused) {              \
        LAUNCH_A2(KS, true);  \
    } else {                  \
        LAUNCH_A2(KS, false); \
    }
    switch (ks) {
        case 1: LAUNCH_A(1); break;
        case 2: LAUNCH_A(2); break;
        case 4: LAUNCH_A(4); break;
        case 8: LAUNCH_A(8); break;
        default: LAUNCH_A(16); break;
    }
#undef LAUNCH_A3
#undef LAUNCH_A2
#undef LAUNCH_A
    CTCB_LAUNCH_CHECK();
    return OK;
}

extern "C" CTCB200_API int ctcb200_ctc_loss_bwd(const float* log_probs, const int64_t* targets, int64_t target_stride,
                                    const int64_t* input_lengths, const int64_t* target_lengths, int T, int N,
                                    int C, int max_target_len, int blank, const float* alpha_ws, const float* nll,
                                    const float* grad_nll, float grad_scale, float* grad, ctcb200_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    CTCB_REQUIRE(T > 0 && N > 0 && C > 0, "ctc_loss_bwd: empty shape T=%d N=%d C=%d", T, N, C);
    CTCB_REQUIRE(blank >= 0 && blank < C, "ctc_loss_bwd: blank %d out of range [0,%d)", blank, C);
    int ks = ks_for(max_target_len);
    CTCB_REQUIRE(ks <= 16, "ctc_loss_bwd: target length %d exceeds the supported maximum 255", max_target_len);
    (void)nll;
    CTCB_REQUIRE((reinterpret_cast<uintptr_t>(alpha_ws) & 15) == 0, "ctc_loss_bwd: workspace must be 16-byte aligned");
    if (ctc_fused(N)) {
        CtcWs w = carve(const_cast<float*>(alpha_ws), T, N, ks);
        const bool rs64 = C <= 64;
        const int C_pad = rs64 ? 64 : ((C + 3) & ~3);
        const size_t smem = static_cast<size_t>(WARPS_PER_BLOCK) * (static_cast<size_t>(PF) * (C_pad + ks * 32) + C_pad) * sizeof(float);
        CTCB_REQUIRE(smem <= 200 * 1024, "ctc_loss_bwd: class count %d too large for the row ring", C);
        dim3 grid((N + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK), block(WARPS_PER_BLOCK * 32);
#define LAUNCH_F2(KS, RS)                                                                                                     \
    CTCB_CUDA(cudaFuncSetAttribute(ctc_beta_grad_kernel<KS, RS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
    ctc_beta_grad_kernel<KS, RS><<<grid, block, smem, stream>>>(log_probs, targets, target_stride, input_lengths,             \
                                                                target_lengths, w.hist_a, w.off_a, w.nll_d, grad_nll,         \
                                                                grad_scale, grad, T, N, C, blank)
#define LAUNCH_F(KS)          \
    if (rs64) {               \
        LAUNCH_F2(KS, 64);    \
    } else {                  \
        LAUNCH_F2(KS, 0);     \
    }
        switch (ks) {
            case 1: LAUNCH_F(1); break;
            case 2: LAUNCH_F(2); break;
            case 4: LAUNCH_F(4); break;
            case 8: LAUNCH_F(8); break;
            default: LAUNCH_F(16); break;
        }
#undef LAUNCH_F2
#undef LAUNCH_F
        CTCB_LAUNCH_CHECK();
        return OK;
    }
    const long long rows = static_cast<long long>(T) * N;
    long long blocks = (rows + 7) / 8;
    const long long cap = static_cast<long long>(device_sm_count()) * 8;
    if (blocks > cap) blocks = cap;
    size_t smem = static_cast<size_t>(8) * C * sizeof(float);
    CTCB_REQUIRE(smem <= 200 * 1024, "ctc_loss_bwd: class count %d too large for the occupancy buffer", C);
    CtcWs w = carve(const_cast<float*>(alpha_ws), T, N, ks);
#define LAUNCH_B(KS)                                                                                               \
    CTCB_CUDA(cudaFuncSetAttribute(ctc_grad_kernel<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
    ctc_grad_kernel<KS><<<static_cast<int>(blocks), 256, smem, stream>>>(log_probs, targets, target_stride, input_lengths,     \
                                                       target_lengths, w.hist_a, w.hist_b, w.off_a, w.off_b, w.nll_d, \
                                                       grad_nll, grad_scale, grad, T, N, C, blank)
    switch (ks) {
        case 1: LAUNCH_B(1); break;
        case 2: LAUNCH_B(2); break;
        case 4: LAUNCH_B(4); break;
        case 8: LAUNCH_B(8); break;
        default: LAUNCH_B(16); break;
    }
#undef LAUNCH_B
    CTCB_LAUNCH_CHECK();
    return OK;
}
