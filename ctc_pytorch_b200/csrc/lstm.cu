// K2 / K3 — persistent recurrent LSTM kernels (forward scan and BPTT) for one bidirectional layer.
//
// Replaces the time loop inside nn.LSTM(bias=False, bidirectional=True) that the reference calls at
// timit/models/model_ctc.py:23-26,33 (forward) and back-propagates through at
// timit/steps/train_ctc.py:63. Semantics reproduced exactly: h0 = c0 = 0, gate order i,f,g,o,
// c_t = sig(f) c_{t-1} + sig(i) tanh(g), h_t = sig(o) tanh(c_t), the reverse direction scans
// t = T-1 .. 0 over *all* padded frames (the reference does not pack sequences).
//
// Decomposition. The input projection x_t W_ih^T for all t is one big tensor-core GEMM (gemm.cu)
// done beforehand; what remains per step is the thin product W_hh h_{t-1} ([4H x H] x [H x N]).
// W_hh never leaves the chip: each CTA owns 32 hidden units = 128 gate rows of one direction,
// keeps that [128 x H] bf16 slice resident in shared memory (loaded once by TMA, SWIZZLE_128B) and
// per step issues one tcgen05.mma chain (M=128, N=batch tile, K=H) into a TMEM accumulator.
// The only per-step traffic is the all-gather of h_t (H x NB bf16) between the H/32 CTAs of a
// (direction, batch-group): every CTA writes its 32 units into a global "operand image" that is
// already laid out as the next step's K-major SWIZZLE_128B B operand, releases a counter, and all
// CTAs copy the image back into shared memory. Gate non-linearities, the cell update and the
// hadamard products are fused in registers between tcgen05.ld and the image store.
//
// The backward kernel mirrors this with W_hh^T: CTA (mb, q) of a 4-CTA cluster holds the
// [128 units x H] slice of gate q's transposed block, multiplies it with gate q's dG image, and the
// four partial dh blocks of a cluster are reduce-scattered through distributed shared memory so
// that each CTA finishes 32 units: dh -> (do, dc, di, df, dg) -> next dG images.
#include <cuda_fp16.h>

#include "common.cuh"
#include "ctcb200.h"

namespace ctcb200 {
namespace {

constexpr int LSTM_THREADS = 256;

__device__ __forceinline__ void wait_counter(const unsigned int* flag, unsigned int target) {
    if (ld_acquire(flag) >= target) return;
    const long long t0 = clock64();
    while (ld_acquire(flag) < target) {
        if (clock64() - t0 > SPIN_LIMIT_CYCLES) spin_timeout_trap(2);
    }
}

__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) { return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * x)); }

// Pack the bf16 of 8 consecutive lanes (this lane = lowest) into a uint4; valid on lanes % 8 == 0.
__device__ __forceinline__ uint4 pack8_bf16(float v) {
    __nv_bfloat16 b = __float2bfloat16(v);
    uint32_t x = static_cast<uint32_t>(*reinterpret_cast<unsigned short*>(&b));
    uint32_t p1 = x | (__shfl_down_sync(0xffffffffu, x, 1) << 16);
    uint32_t p2 = __shfl_down_sync(0xffffffffu, p1, 2);
    uint32_t p3 = __shfl_down_sync(0xffffffffu, p1, 4);
    uint32_t p4 = __shfl_down_sync(0xffffffffu, p2, 4);
    return make_uint4(p1, p2, p3, p4);
}

// Element offset (bf16 units) of unit `u`, batch row `n` inside a [H x NB] K-major SWIZZLE_128B image.
template <int NB>
__device__ __forceinline__ int image_chunk_offset(int u, int n) {
    const int kb = u >> 6, c = (u & 63) >> 3;
    return kb * (NB * 64) + n * 64 + ((c ^ (n & 7)) << 3);
}

struct FwdParams {
    const float* gx;          // [T*N, 8H] gate pre-activations from the input projection (packed column order)
    float* hout;              // [T*N, 2H] layer output (fwd | reverse)
    float* c_save;            // [T*N, 2H] cell states, or null (inference)
    uint2* gates_save;        // [T*N, 2H] activated gates as 4 x fp16 (i,f,g,o), or null
    __nv_bfloat16* himg;      // [2 dirs][groups][2][H*NB] operand images
    unsigned int* flags;      // [2 dirs][groups] step counters, 32 uints apart
    int T, N, H, groups, n0;  // n0 = first batch row of this launch's group 0
};

template <int NB>
__global__ void __launch_bounds__(LSTM_THREADS, 1)
lstm_fwd_kernel(const __grid_constant__ CUtensorMap tmW, FwdParams p) {
    constexpr int CPT = NB / 2;        // accumulator columns per thread
    constexpr int EPT = NB / 8;        // (unit, batch) elements per thread in the cell update
    constexpr int S_STRIDE = NB * 4 + 4;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int H = p.H, T = p.T, N = p.N;
    uint8_t* sW = smem;
    uint8_t* sH = sW + 128 * H * 2;
    float* sS = reinterpret_cast<float*>(sH + H * NB * 2);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sS + 32 * S_STRIDE);
    uint64_t* w_full = bars;
    uint64_t* h_full = bars + 1;
    uint64_t* acc_full = bars + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int j = blockIdx.x, dir = blockIdx.y, grp = blockIdx.z;
    const int ctas = gridDim.x;
    const int kblocks = H / 64;

    if (tid == 0) {
        tma_prefetch_desc(&tmW);
        mbar_init(w_full, 1);
        mbar_init(h_full, LSTM_THREADS);
        mbar_init(acc_full, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, NB < 32 ? 32 : NB);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (tid == 0) {
        mbar_expect_tx(w_full, 128 * H * 2);
        for (int kb = 0; kb < kblocks; ++kb) tma_load_2d(sW + kb * 16384, &tmW, w_full, kb * 64, dir * 4 * H + j * 128);
    }

    const int lq = warp & 3, ch = warp >> 2;
    const int row = lq * 32 + lane;            // gate row inside this CTA's 128
    const int u_loc = row >> 2, q = row & 3;   // hidden unit (0..31) and gate (i,f,g,o)
    const float act_s = (q == 2) ? 2.0f : 1.0f;
    const size_t gx_col = static_cast<size_t>(dir) * 4 * H + j * 128 + row;
    const size_t G8 = static_cast<size_t>(8) * H, H2 = static_cast<size_t>(2) * H;
    __nv_bfloat16* img = p.himg + (static_cast<size_t>(dir) * p.groups + grp) * 2 * H * NB;
    unsigned int* flag = p.flags + (dir * p.groups + grp) * 32;
    const int chunks = H * NB / 8;
    constexpr uint32_t idesc = umma_idesc_bf16(128, NB);

    float c_state[EPT];
#pragma unroll
    for (int e = 0; e < EPT; ++e) c_state[e] = 0.0f;

    for (int t = 0; t < T; ++t) {
        const int tt = dir ? (T - 1 - t) : t;
        // (1) this step's input-projection terms: independent of the recurrence, issued first
        float gx[CPT];
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
            const int gn = p.n0 + grp * NB + ch * CPT + c;
            gx[c] = (gn < N) ? __ldg(p.gx + (static_cast<size_t>(tt) * N + gn) * G8 + gx_col) : 0.0f;
        }
        // (2) all CTAs of this (direction, group) have published h_{t-1}
        if (lane == 0) wait_counter(flag, static_cast<unsigned int>(ctas) * t);
        __syncwarp();
        // (3) operand image -> shared memory
        {
            const uint4* src = reinterpret_cast<const uint4*>(img + static_cast<size_t>(t & 1) * H * NB);
            uint4* dst = reinterpret_cast<uint4*>(sH);
            for (int i = tid; i < chunks; i += LSTM_THREADS) dst[i] = ld_cg_v4(src + i);
            fence_proxy_async_smem();
            mbar_arrive(h_full);
        }
        // (4) one thread issues the K = H MMA chain
        if (tid == 0) {
            if (t == 0) mbar_wait(w_full, 0);
            mbar_wait(h_full, t & 1);
            tc_fence_after();
            const uint32_t a0 = smem_u32(sW), b0 = smem_u32(sH);
            for (int kb = 0; kb < kblocks; ++kb) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    umma_bf16(tmem_base, umma_desc_sw128(a0 + kb * 16384 + kk * 32),
                              umma_desc_sw128(b0 + kb * (NB * 128) + kk * 32), idesc, (kb | kk) != 0 ? 1u : 0u);
            }
            umma_commit(acc_full);
        }
        __syncwarp();
        // (5) accumulator -> registers
        mbar_wait(acc_full, t & 1);
        tc_fence_after();
        uint32_t acc[CPT];
        if constexpr (CPT == 16) tmem_ld_32x16(tmem_base + (static_cast<uint32_t>(lq * 32) << 16) + ch * CPT, acc);
        else tmem_ld_32x8(tmem_base + (static_cast<uint32_t>(lq * 32) << 16) + ch * CPT, acc);
        tmem_ld_wait();
        tc_fence_before();
        // (6) gate non-linearity, then regroup the four gates of a unit through shared memory
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
            const float pre = __uint_as_float(acc[c]) + gx[c];
            const float a = act_s * fast_sigmoid(act_s * pre) - (act_s - 1.0f);  // sigmoid, or tanh for gate g
            sS[u_loc * S_STRIDE + (ch * CPT + c) * 4 + q] = a;
        }
        __syncthreads();
        // (7) cell update for (unit = lane, batch n = warp + 8e); publish h_t into the next image
        float hv[EPT];
        float4 gv[EPT];
        __nv_bfloat16* img_next = img + static_cast<size_t>((t + 1) & 1) * H * NB;
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
            const int n = warp + 8 * e;
            const float4 g4 = *reinterpret_cast<const float4*>(&sS[lane * S_STRIDE + n * 4]);
            const float cn = g4.y * c_state[e] + g4.x * g4.z;
            c_state[e] = cn;
            const float h = g4.w * fast_tanh(cn);
            hv[e] = h;
            gv[e] = g4;
            const uint4 pk = pack8_bf16(h);
            if ((lane & 7) == 0)
                *reinterpret_cast<uint4*>(img_next + image_chunk_offset<NB>(j * 32 + lane, n)) = pk;
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) red_release_add(flag, 1u);
        // (8) off the critical path: layer output and the activations BPTT needs
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
            const int gn = p.n0 + grp * NB + warp + 8 * e;
            if (gn < N) {
                const size_t o = (static_cast<size_t>(tt) * N + gn) * H2 + static_cast<size_t>(dir) * H + j * 32 + lane;
                p.hout[o] = hv[e];
                if (p.c_save) p.c_save[o] = c_state[e];
                if (p.gates_save) {
                    __half2 lo = __floats2half2_rn(gv[e].x, gv[e].y), hi = __floats2half2_rn(gv[e].z, gv[e].w);
                    p.gates_save[o] = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, NB < 32 ? 32 : NB);
}

// ------------------------------------------------------------------------------------------------
// backward (BPTT)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mapa_shared(uint32_t local_addr, uint32_t cta_rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta_rank));
    return r;
}
__device__ __forceinline__ void st_cluster_f32(uint32_t addr, float v) {
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

struct BwdParams {
    const float* dhout;        // [T*N, 2H] gradient w.r.t. the layer output
    const float* c_save;       // [T*N, 2H]
    const uint2* gates_save;   // [T*N, 2H] 4 x fp16
    __nv_bfloat16* dg;         // [T*N, 8H] gate gradients, packed column order (A operand of the dX GEMM)
    __nv_bfloat16* dgimg;      // [2 dirs][groups][4 gates][2][H*NB] operand images
    unsigned int* flags;       // [2 dirs][groups]
    int T, N, H, groups, n0;
};

template <int NB>
__global__ void __launch_bounds__(LSTM_THREADS, 1)
lstm_bwd_kernel(const __grid_constant__ CUtensorMap tmWT, BwdParams p) {
    constexpr int CPT = NB / 2;
    constexpr int EPT = NB / 8;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int H = p.H, T = p.T, N = p.N;
    uint8_t* sW = smem;
    uint8_t* sB = sW + 128 * H * 2;
    float* sR = reinterpret_cast<float*>(sB + H * NB * 2);  // [4 src][NB][32] partial dh blocks
    uint64_t* bars = reinterpret_cast<uint64_t*>(sR + 4 * NB * 32);
    uint64_t* w_full = bars;
    uint64_t* b_full = bars + 1;
    uint64_t* acc_full = bars + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = blockIdx.x;  // gate handled by this CTA == rank in the 4-CTA cluster
    const int mb = blockIdx.y;
    const int dir = blockIdx.z / p.groups, grp = blockIdx.z % p.groups;
    const int ctas = 4 * gridDim.y;
    const int kblocks = H / 64;

    if (tid == 0) {
        tma_prefetch_desc(&tmWT);
        mbar_init(w_full, 1);
        mbar_init(b_full, LSTM_THREADS);
        mbar_init(acc_full, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, NB < 32 ? 32 : NB);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    cluster_sync_all();  // every CTA of the cluster has its barriers / receive buffer ready

    if (tid == 0) {
        mbar_expect_tx(w_full, 128 * H * 2);
        // rows of the transposed gate block: (dir, q, unit); this CTA takes units [128 mb, +128)
        for (int kb = 0; kb < kblocks; ++kb)
            tma_load_2d(sW + kb * 16384, &tmWT, w_full, kb * 64, (dir * 4 + q) * H + mb * 128);
    }

    const int lq = warp & 3, ch = warp >> 2;
    const int unit = mb * 128 + q * 32 + lane;  // the unit this thread finishes in the element phase
    const size_t H2 = static_cast<size_t>(2) * H, G8 = static_cast<size_t>(8) * H;
    const size_t dg_col = static_cast<size_t>(dir) * 4 * H + static_cast<size_t>(unit >> 5) * 128 + (unit & 31) * 4;
    __nv_bfloat16* imgs = p.dgimg + (static_cast<size_t>(dir) * p.groups + grp) * 4 * 2 * H * NB;
    unsigned int* flag = p.flags + (dir * p.groups + grp) * 32;
    const int chunks = H * NB / 8;
    constexpr uint32_t idesc = umma_idesc_bf16(128, NB);
    // remote receive slot: partial block from source gate q lands in CTA `lq` (owner of rows 32 lq..)
    const uint32_t remote_base = mapa_shared(smem_u32(sR + (q * NB + ch * CPT) * 32 + lane), static_cast<uint32_t>(lq));

    float dc_carry[EPT];
#pragma unroll
    for (int e = 0; e < EPT; ++e) dc_carry[e] = 0.0f;

    for (int t = 0; t < T; ++t) {
        const int tt = dir ? t : (T - 1 - t);          // reverse of the forward scan order
        const int tprev = dir ? tt + 1 : tt - 1;       // time index that held c_{prev} in the forward scan
        const bool has_prev = dir ? (tt + 1 < T) : (tt >= 1);
        // (1) saved activations and the incoming gradient for this thread's elements
        float dh_in[EPT], c_t[EPT], c_p[EPT];
        uint2 gts[EPT];
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
            const int gn = p.n0 + grp * NB + warp + 8 * e;
            const bool ok = gn < N;
            const size_t o = (static_cast<size_t>(tt) * N + (ok ? gn : 0)) * H2 + static_cast<size_t>(dir) * H + unit;
            dh_in[e] = ok ? __ldg(p.dhout + o) : 0.0f;
            c_t[e] = ok ? __ldg(p.c_save + o) : 0.0f;
            gts[e] = ok ? __ldg(p.gates_save + o) : make_uint2(0u, 0u);
            const size_t op = (static_cast<size_t>(has_prev ? tprev : tt) * N + (ok ? gn : 0)) * H2 +
                              static_cast<size_t>(dir) * H + unit;
            c_p[e] = (ok && has_prev) ? __ldg(p.c_save + op) : 0.0f;
        }
        // (2) gate gradients of the previous BPTT step are published
        if (lane == 0) wait_counter(flag, static_cast<unsigned int>(ctas) * t);
        __syncwarp();
        // (3) dG image of gate q -> shared memory
        {
            const uint4* src = reinterpret_cast<const uint4*>(imgs + (static_cast<size_t>(q) * 2 + (t & 1)) * H * NB);
            uint4* dst = reinterpret_cast<uint4*>(sB);
            for (int i = tid; i < chunks; i += LSTM_THREADS) dst[i] = ld_cg_v4(src + i);
            fence_proxy_async_smem();
            mbar_arrive(b_full);
        }
        // (4) partial dh[128 units, NB] = W_q^T slice * dG_q
        if (tid == 0) {
            if (t == 0) mbar_wait(w_full, 0);
            mbar_wait(b_full, t & 1);
            tc_fence_after();
            const uint32_t a0 = smem_u32(sW), b0 = smem_u32(sB);
            for (int kb = 0; kb < kblocks; ++kb) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    umma_bf16(tmem_base, umma_desc_sw128(a0 + kb * 16384 + kk * 32),
                              umma_desc_sw128(b0 + kb * (NB * 128) + kk * 32), idesc, (kb | kk) != 0 ? 1u : 0u);
            }
            umma_commit(acc_full);
        }
        __syncwarp();
        // (5) scatter the partial rows to their owner CTA through distributed shared memory
        mbar_wait(acc_full, t & 1);
        tc_fence_after();
        uint32_t acc[CPT];
        if constexpr (CPT == 16) tmem_ld_32x16(tmem_base + (static_cast<uint32_t>(lq * 32) << 16) + ch * CPT, acc);
        else tmem_ld_32x8(tmem_base + (static_cast<uint32_t>(lq * 32) << 16) + ch * CPT, acc);
        tmem_ld_wait();
        tc_fence_before();
#pragma unroll
        for (int c = 0; c < CPT; ++c) st_cluster_f32(remote_base + c * 32 * 4, __uint_as_float(acc[c]));
        cluster_sync_all();
        // (6) finish 32 units: recurrent dh, LSTM cell backward, publish the four gate gradients
        uint2 dgp[EPT];
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
            const int n = warp + 8 * e;
            float dh = dh_in[e];
#pragma unroll
            for (int s = 0; s < 4; ++s) dh += sR[(s * NB + n) * 32 + lane];
            const __half2 lo = *reinterpret_cast<const __half2*>(&gts[e].x);
            const __half2 hi = *reinterpret_cast<const __half2*>(&gts[e].y);
            const float gi = __low2float(lo), gf = __high2float(lo), gg = __low2float(hi), go = __high2float(hi);
            const float tc = fast_tanh(c_t[e]);
            const float d_o = dh * tc * go * (1.0f - go);
            const float dc = dc_carry[e] + dh * go * (1.0f - tc * tc);
            const float d_i = dc * gg * gi * (1.0f - gi);
            const float d_f = dc * c_p[e] * gf * (1.0f - gf);
            const float d_g = dc * gi * (1.0f - gg * gg);
            dc_carry[e] = dc * gf;
            const int off = image_chunk_offset<NB>(unit, n);
            const size_t nxt = static_cast<size_t>((t + 1) & 1) * H * NB;
            const uint4 pi = pack8_bf16(d_i), pf = pack8_bf16(d_f), pg = pack8_bf16(d_g), po = pack8_bf16(d_o);
            if ((lane & 7) == 0) {
                *reinterpret_cast<uint4*>(imgs + (0 * 2) * static_cast<size_t>(H) * NB + nxt + off) = pi;
                *reinterpret_cast<uint4*>(imgs + (1 * 2) * static_cast<size_t>(H) * NB + nxt + off) = pf;
                *reinterpret_cast<uint4*>(imgs + (2 * 2) * static_cast<size_t>(H) * NB + nxt + off) = pg;
                *reinterpret_cast<uint4*>(imgs + (3 * 2) * static_cast<size_t>(H) * NB + nxt + off) = po;
            }
            __nv_bfloat162 b01 = __floats2bfloat162_rn(d_i, d_f), b23 = __floats2bfloat162_rn(d_g, d_o);
            dgp[e] = make_uint2(*reinterpret_cast<uint32_t*>(&b01), *reinterpret_cast<uint32_t*>(&b23));
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) red_release_add(flag, 1u);
        // (7) off the critical path: dG rows for the dX / dW GEMMs
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
            const int gn = p.n0 + grp * NB + warp + 8 * e;
            if (gn < N)
                *reinterpret_cast<uint2*>(p.dg + (static_cast<size_t>(tt) * N + gn) * G8 + dg_col) = dgp[e];
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // no CTA exits while a peer may still address its shared memory
    if (warp == 1) tmem_dealloc(tmem_base, NB < 32 ? 32 : NB);
}

template <int NB>
size_t lstm_smem_bytes(int H, bool bwd) {
    size_t b = static_cast<size_t>(128) * H * 2 + static_cast<size_t>(H) * NB * 2;
    b += bwd ? static_cast<size_t>(4) * NB * 32 * 4 : static_cast<size_t>(32) * (NB * 4 + 4) * 4;
    return b + 64 + 1024;
}

int pick_nb(int N, int H, int force_nb, bool bwd) {
    if (force_nb == 16 || force_nb == 32) return force_nb;
    const int sms = device_sm_count();
    const int per_group = bwd ? 2 * 4 * (H / 128) : 2 * (H / 32);
    // prefer the narrow batch tile (shorter per-step chain) when every group fits at once
    const int g16 = (N + 15) / 16;
    const int budget = bwd ? (sms / 4) * 4 - 16 : sms;  // clusters of 4 cannot use every SM
    if (g16 * per_group <= budget) return 16;
    return 32;
}

}  // namespace
}  // namespace ctcb200

using namespace ctcb200;

extern "C" CTCB200_API int64_t ctcb200_lstm_scratch_bytes(int N, int H) {
    // two directions x groups(16-wide worst case) x (4 gates x 2 parities) images + flags
    const int64_t groups = (N + 15) / 16;
    return 2 * groups * 8 * static_cast<int64_t>(H) * 32 * 2 + 2 * groups * 32 * 4 + 1024;
}

extern "C" CTCB200_API int ctcb200_lstm_fwd(const float* gx, const void* whh_packed, float* hout, float* c_save,
                                            void* gates_save, void* scratch, int T, int N, int H, int batch_tile,
                                            ctcb200_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    CTCB_REQUIRE(T > 0 && N > 0, "lstm_fwd: empty T=%d N=%d", T, N);
    CTCB_REQUIRE(H % 128 == 0 && H >= 128 && H <= 640, "lstm_fwd: hidden size %d must be a multiple of 128 in [128,640]", H);
    const int NB = pick_nb(N, H, batch_tile, false);
    const int groups_total = (N + NB - 1) / NB;
    const int per_group = 2 * (H / 32);
    const int sms = device_sm_count();
    CTCB_REQUIRE(per_group <= sms, "lstm_fwd: one batch group needs %d CTAs but the device has %d SMs", per_group, sms);
    const int groups_per_launch = sms / per_group;
    CUtensorMap tmW;
    int rc = make_tmap_bf16_2d(&tmW, whh_packed, static_cast<uint64_t>(8) * H, H, H, 128, 64);
    if (rc != OK) return rc;
    const size_t smem = NB == 16 ? lstm_smem_bytes<16>(H, false) : lstm_smem_bytes<32>(H, false);
    CTCB_REQUIRE(smem <= 227 * 1024, "lstm_fwd: shared memory %zu exceeds 227 KB (H=%d)", smem, H);
    void* kern = NB == 16 ? reinterpret_cast<void*>(lstm_fwd_kernel<16>) : reinterpret_cast<void*>(lstm_fwd_kernel<32>);
    CTCB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    uint8_t* scr = static_cast<uint8_t*>(scratch);
    for (int g0 = 0; g0 < groups_total; g0 += groups_per_launch) {
        const int groups = (groups_total - g0 < groups_per_launch) ? groups_total - g0 : groups_per_launch;
        const size_t img_bytes = static_cast<size_t>(2) * groups * 2 * H * NB * 2;
        const size_t flag_bytes = static_cast<size_t>(2) * groups * 32 * 4;
        CTCB_CUDA(cudaMemsetAsync(scr, 0, img_bytes + flag_bytes, stream));
        FwdParams p;
        p.gx = gx; p.hout = hout; p.c_save = c_save; p.gates_save = static_cast<uint2*>(gates_save);
        p.himg = reinterpret_cast<__nv_bfloat16*>(scr);
        p.flags = reinterpret_cast<unsigned int*>(scr + img_bytes);
        p.T = T; p.N = N; p.H = H; p.groups = groups; p.n0 = g0 * NB;
        void* args[] = {const_cast<CUtensorMap*>(&tmW), &p};
        dim3 grid(H / 32, 2, groups), block(LSTM_THREADS);
        CTCB_CUDA(cudaLaunchCooperativeKernel(kern, grid, block, args, smem, stream));
    }
    return OK;
}

extern "C" CTCB200_API int ctcb200_lstm_bwd(const float* dhout, const void* whhT_packed, const float* c_save,
                                            const void* gates_save, void* dg, void* scratch, int T, int N, int H,
                                            int batch_tile, ctcb200_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    CTCB_REQUIRE(T > 0 && N > 0, "lstm_bwd: empty T=%d N=%d", T, N);
    CTCB_REQUIRE(H % 128 == 0 && H >= 128 && H <= 640, "lstm_bwd: hidden size %d must be a multiple of 128 in [128,640]", H);
    const int NB = pick_nb(N, H, batch_tile, true);
    const int groups_total = (N + NB - 1) / NB;
    const int per_group = 2 * 4 * (H / 128);
    const int sms = device_sm_count();
    const int budget = (sms / 4) * 4 - 16;  // clusters of 4 strand a few SMs
    CTCB_REQUIRE(per_group <= budget, "lstm_bwd: one batch group needs %d CTAs; device budget %d", per_group, budget);
    const int groups_per_launch = budget / per_group;
    CUtensorMap tmWT;
    int rc = make_tmap_bf16_2d(&tmWT, whhT_packed, static_cast<uint64_t>(8) * H, H, H, 128, 64);
    if (rc != OK) return rc;
    const size_t smem = NB == 16 ? lstm_smem_bytes<16>(H, true) : lstm_smem_bytes<32>(H, true);
    CTCB_REQUIRE(smem <= 227 * 1024, "lstm_bwd: shared memory %zu exceeds 227 KB (H=%d)", smem, H);
    auto kern = NB == 16 ? lstm_bwd_kernel<16> : lstm_bwd_kernel<32>;
    CTCB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    uint8_t* scr = static_cast<uint8_t*>(scratch);
    for (int g0 = 0; g0 < groups_total; g0 += groups_per_launch) {
        const int groups = (groups_total - g0 < groups_per_launch) ? groups_total - g0 : groups_per_launch;
        const size_t img_bytes = static_cast<size_t>(2) * groups * 4 * 2 * H * NB * 2;
        const size_t flag_bytes = static_cast<size_t>(2) * groups * 32 * 4;
        CTCB_CUDA(cudaMemsetAsync(scr, 0, img_bytes + flag_bytes, stream));
        BwdParams p;
        p.dhout = dhout; p.c_save = c_save; p.gates_save = static_cast<const uint2*>(gates_save);
        p.dg = static_cast<__nv_bfloat16*>(dg);
        p.dgimg = reinterpret_cast<__nv_bfloat16*>(scr);
        p.flags = reinterpret_cast<unsigned int*>(scr + img_bytes);
        p.T = T; p.N = N; p.H = H; p.groups = groups; p.n0 = g0 * NB;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(4, H / 128, 2 * groups);
        cfg.blockDim = dim3(LSTM_THREADS);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = stream;
        cudaLaunchAttribute attrs[2];
        attrs[0].id = cudaLaunchAttributeClusterDimension;
        attrs[0].val.clusterDim.x = 4; attrs[0].val.clusterDim.y = 1; attrs[0].val.clusterDim.z = 1;
        attrs[1].id = cudaLaunchAttributeCooperative;
        attrs[1].val.cooperative = 1;
        cfg.attrs = attrs;
        cfg.numAttrs = 2;
        cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tmWT, p);
        if (e != cudaSuccess) {
            // some drivers refuse cooperative + cluster together; co-residency is then guaranteed by the CTA
            // budget above (one CTA per SM, grid <= schedulable clusters) on an otherwise idle device
            (void)cudaGetLastError();
            cfg.numAttrs = 1;
            e = cudaLaunchKernelEx(&cfg, kern, tmWT, p);
        }
        CTCB_CUDA(e);
    }
    return OK;
}
