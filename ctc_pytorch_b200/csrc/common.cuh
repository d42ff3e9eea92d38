// Shared device/host helpers for the sm_100a kernels of libctcb200.
//
// Everything here is a thin wrapper over one PTX instruction (mbarrier, TMA, tcgen05, TMEM) or a
// host-side utility (error slot, tensor-map encoding through the runtime's driver entry point so
// the library has no link-time dependency on libcuda).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

// ---------------------------------------------------------------------------------------------
// error slot (host)
// ---------------------------------------------------------------------------------------------
namespace ctcb200 {

enum Status : int {
    OK = 0,
    ERR_INVALID = -1,    // bad argument / unsupported shape
    ERR_CUDA = -2,       // a CUDA runtime call failed
    ERR_DRIVER = -3,     // driver entry point (tensor map encode) failed
    ERR_TIMEOUT = -4,    // device-side spin wait gave up (would have dead-locked)
};

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define CTCB_CUDA(expr)                                                              \
    do {                                                                             \
        cudaError_t _e = (expr);                                                     \
        if (_e != cudaSuccess) return ::ctcb200::cuda_fail(_e, #expr, __FILE__, __LINE__); \
    } while (0)

#define CTCB_REQUIRE(cond, ...)                  \
    do {                                         \
        if (!(cond)) {                           \
            ::ctcb200::set_error(__VA_ARGS__);   \
            return ::ctcb200::ERR_INVALID;       \
        }                                        \
    } while (0)

#define CTCB_LAUNCH_CHECK() CTCB_CUDA(cudaGetLastError())

// Encode a 2-D bf16 tensor map: global tensor [rows][cols] (cols contiguous, row pitch
// `row_stride_elems` elements), box [box_rows][box_cols], SWIZZLE_128B when box_cols*2 == 128.
int make_tmap_bf16_2d(CUtensorMap* out, const void* gptr, uint64_t rows, uint64_t cols,
                      uint64_t row_stride_elems, uint32_t box_rows, uint32_t box_cols);

int make_tmap_f32_2d(CUtensorMap* out, const void* gptr, uint64_t rows, uint64_t cols, uint64_t row_stride_elems,
                     uint32_t box_rows, uint32_t box_cols);

// Per-device host-side state (function attributes, probe caches, SM counts) is indexed by the CUDA device ordinal.
constexpr int MAX_DEVICES = 64;
int current_device();
int gemm_preload();   // gemm.cu: load every GEMM tile width on the current device (see gemm_prepare)
int device_sm_count();

}  // namespace ctcb200

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__
namespace ctcb200 {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(pred));
    return pred != 0;
}

// ---- mbarrier --------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Watchdog: a wait that lasts longer than ~2 s of SM clocks means a protocol bug (lost TMA, wrong
// parity); trap so the launch fails with an error instead of hanging the device.
constexpr long long SPIN_LIMIT_CYCLES = 4000000000LL;
static __device__ __noinline__ void spin_timeout_trap(int what) {
    printf("ctcb200: device wait timed out (kind %d) block %d thread %d\n", what, blockIdx.x, threadIdx.x);
    __trap();
}
// Tagged variant for kernels with several roles: on a timeout every waiting warp reports WHICH hand-off it was waiting for and
// at which step before the launch is trapped (all warps get ~0.5 s to report), so a protocol stall can be read off the log.
static __device__ __noinline__ void spin_timeout_report(int tag, int step) {
    if ((threadIdx.x & 31) == 0)
        printf("ctcb200: wait timed out: tag %d step %d block (%d,%d,%d) warp %d\n", tag, step, blockIdx.x, blockIdx.y, blockIdx.z,
               threadIdx.x >> 5);
}
#ifdef CTCB200_WAIT_TAGS   // debugging build: -DCTCB200_WAIT_TAGS (python -m ctc_pytorch_b200._build --tags)
__device__ __forceinline__ void mbar_wait_tag(uint64_t* bar, uint32_t parity, int tag, int step) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    bool reported = false;
    while (!mbar_try_wait(bar, parity)) {
        const long long dt = clock64() - t0;
        if (dt > SPIN_LIMIT_CYCLES && !reported) { spin_timeout_report(tag, step); reported = true; }
        if (dt > SPIN_LIMIT_CYCLES + SPIN_LIMIT_CYCLES / 4) __trap();
    }
}
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > SPIN_LIMIT_CYCLES) spin_timeout_trap(1);
    }
}
#ifndef CTCB200_WAIT_TAGS
__device__ __forceinline__ void mbar_wait_tag(uint64_t* bar, uint32_t parity, int, int) { mbar_wait(bar, parity); }
#endif

// ---- proxies / fences ------------------------------------------------------------------------
// generic-proxy writes to shared memory -> visible to the async proxy (TMA / tcgen05.mma reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// ---- TMA -------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}
// 2-D tile load global -> shared, completion counted in bytes on `bar`. c0 = inner (contiguous)
// coordinate, c1 = outer coordinate, both in elements.
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tm, uint64_t* bar, int c0,
                                            int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

// 2-D tile store shared -> global (clipped to the tensor bounds by the hardware), tracked by bulk groups
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, const void* smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(tm)),
                 "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
                 : "memory");
}
// same, but the tile is ADDED to global memory (fp32 reduction performed by the TMA unit / L2): split-K partials
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* tm, const void* smem_src, int c0, int c1) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(tm)),
                 "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void bulk_group_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_group_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_group_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---- TMEM / tcgen05 --------------------------------------------------------------------------
// One full warp allocates `ncols` (power of two >= 32) TMEM columns; address lands in *smem_slot.
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; kind::f16 covers bf16/fp16 inputs with fp32 accumulate.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Same with the A operand resident in TMEM (128 lanes = rows, 16 K-elements = 8 packed 32-bit columns).
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// registers -> TMEM: this warp's 32 lanes x 8 consecutive 32-bit columns (one row per thread)
__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// Arrive on `bar` when every tcgen05.mma issued so far by this thread has completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// Shared-memory matrix descriptor, K-major operand, SWIZZLE_128B: rows of 128 bytes (64 bf16),
// 8-row swizzle atoms 1024 bytes apart (SBO), tile base 1024-byte aligned.
// Field layout follows the PTX ISA "shared memory descriptor" (sm_100): [0,14) addr>>4,
// [16,30) LBO>>4, [32,46) SBO>>4, [46,48) version=1, [61,64) layout (2 = 128B swizzle).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(1) << 16;            // LBO (unused for swizzled K-major)
    d |= static_cast<uint64_t>(1024 >> 4) << 32;    // SBO
    d |= static_cast<uint64_t>(1) << 46;            // descriptor version (Blackwell)
    d |= static_cast<uint64_t>(2) << 61;            // SWIZZLE_128B
    return d;
}

// MN-major operand (element (mn, k) stored with mn contiguous), SWIZZLE_128B: the tile is a row of 8 KB boxes, each 64 K rows x
// 64 mn elements (128-byte rows, exactly what a TMA box load writes); the canonical layout is ((8,n),(8,k)):((1,LBO),(8,SBO))
// in 16-byte units — 8 chunks of a 128-byte row, then the next 64-wide mn group LBO = 8 KB further; 8 K rows 128 bytes apart,
// then the next group of 8 K rows SBO = 1 KB further. One tcgen05.mma (K = 16) reads two such groups.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(8192 >> 4) << 16;    // LBO: next 64-element group along M / N
    d |= static_cast<uint64_t>(1024 >> 4) << 32;    // SBO: next group of 8 K rows
    d |= static_cast<uint64_t>(1) << 46;            // descriptor version (Blackwell)
    d |= static_cast<uint64_t>(2) << 61;            // SWIZZLE_128B
    return d;
}

// Same for SWIZZLE_64B: rows of 64 bytes (32 bf16), 8-row atoms 512 bytes apart, tile base 512-byte aligned.
// Used for the recurrent kernels' B operand so that the 32 hidden units one CTA produces form one contiguous block.
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>(512 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(4) << 61;            // SWIZZLE_64B
    return d;
}

// Instruction descriptor for kind::f16: bf16 x bf16 -> fp32, both operands K-major, M x N tile.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
    return (1u << 4)                              // D format f32
           | (1u << 7)                            // A format bf16
           | (1u << 10)                           // B format bf16
           | (static_cast<uint32_t>(N >> 3) << 17)
           | (static_cast<uint32_t>(M >> 4) << 24);
}

// TMEM -> registers: this warp's 32 lanes x 32 consecutive fp32 columns (one row per thread).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_32x4(uint32_t taddr, uint32_t (&r)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- global-memory flags (inter-CTA hand-off inside the persistent recurrent kernels) ---------
__device__ __forceinline__ void red_release_add(unsigned int* p, unsigned int v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {   // words published by stream memory operations
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ld_cg_f32(const float* p) {
    float v;
    asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint4 ld_cg_v4(const void* p) {
    uint4 v;
    asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- math --------------------------------------------------------------------------------------
__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_acc(float x) {
    // 1 - 2/(1+e^{2x}); exact limits at +-inf, ~1e-7 abs error elsewhere
    float e = __expf(2.0f * x);
    return 1.0f - 2.0f / (1.0f + e);
}

}  // namespace ctcb200
#endif  // __CUDACC__
