// Library-wide plumbing of the C ABI: thread-local error slot, CUDA error mapping, TMA tensor-map
// encoding via the runtime's driver entry point (no link-time libcuda dependency, so the shared
// object also loads on a machine without a GPU driver — the symbol-export test relies on that).
#include <cstdarg>
#include <cstring>

#include "common.cuh"
#include "ctcb200.h"

namespace ctcb200 {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    set_error("CUDA error %d (%s) at %s:%d: %s", static_cast<int>(e), cudaGetErrorString(e), file, line, what);
    return ERR_CUDA;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
    return fn;
}

int make_tmap_bf16_2d(CUtensorMap* out, const void* gptr, uint64_t rows, uint64_t cols, uint64_t row_stride_elems,
                      uint32_t box_rows, uint32_t box_cols) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled driver entry point unavailable");
        return ERR_DRIVER;
    }
    if ((reinterpret_cast<uintptr_t>(gptr) & 15) != 0 || (row_stride_elems * 2) % 16 != 0) {
        set_error("tensor map: base %p / row pitch %llu B must be 16-byte aligned", gptr,
                  (unsigned long long)(row_stride_elems * 2));
        return ERR_INVALID;
    }
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {row_stride_elems * 2};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUtensorMapSwizzle sw = (box_cols * 2 == 128)  ? CU_TENSOR_MAP_SWIZZLE_128B
                            : (box_cols * 2 == 64) ? CU_TENSOR_MAP_SWIZZLE_64B
                            : (box_cols * 2 == 32) ? CU_TENSOR_MAP_SWIZZLE_32B
                                                   : CU_TENSOR_MAP_SWIZZLE_NONE;
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(gptr), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d): rows=%llu cols=%llu pitch=%llu box=%ux%u", (int)r,
                  (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)row_stride_elems, box_rows,
                  box_cols);
        return ERR_DRIVER;
    }
    return OK;
}

int make_tmap_f32_2d(CUtensorMap* out, const void* gptr, uint64_t rows, uint64_t cols, uint64_t row_stride_elems,
                     uint32_t box_rows, uint32_t box_cols) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled driver entry point unavailable");
        return ERR_DRIVER;
    }
    if ((reinterpret_cast<uintptr_t>(gptr) & 15) != 0 || (row_stride_elems * 4) % 16 != 0) {
        set_error("tensor map (f32): base / row pitch must be 16-byte aligned");
        return ERR_INVALID;
    }
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {row_stride_elems * 4};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUtensorMapSwizzle sw = (box_cols * 4 == 128) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE;
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(gptr), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (f32) failed (%d)", (int)r);
        return ERR_DRIVER;
    }
    return OK;
}

int current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    return (dev >= 0 && dev < MAX_DEVICES) ? dev : 0;
}

int device_sm_count() {
    static int sms[MAX_DEVICES] = {0};   // per device ordinal (a process may drive several GPUs)
    const int dev = current_device();
    if (sms[dev]) return sms[dev];
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) return 148;
    sms[dev] = v;
    return v;
}

}  // namespace ctcb200

extern "C" CTCB200_API const char* ctcb200_last_error(void) { return ctcb200::g_err; }

extern "C" CTCB200_API int ctcb200_version(void) { return 100; }

// Stream-ordered wait on a device word (driver stream memory operation): work enqueued on `stream` after this call
// does not start before *counter >= value (wrap-around compare). No SM is occupied while waiting.
extern "C" CTCB200_API int ctcb200_stream_wait_geq(ctcb200_stream_t stream_, const void* counter, uint32_t value) {
    typedef CUresult (*WaitFn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
    static WaitFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &q);
        if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
            (void)cudaGetLastError();
            ctcb200::set_error("cuStreamWaitValue32 driver entry point unavailable");
            return ctcb200::ERR_DRIVER;
        }
        fn = reinterpret_cast<WaitFn>(p);
    }
    if (!counter || (reinterpret_cast<uintptr_t>(counter) & 3) != 0) {
        ctcb200::set_error("stream_wait_geq: counter %p must be a 4-byte aligned device address", counter);
        return ctcb200::ERR_INVALID;
    }
    CUresult r = fn(static_cast<CUstream>(stream_), reinterpret_cast<CUdeviceptr>(counter), value, 0u /* GEQ */);
    if (r != CUDA_SUCCESS) {
        ctcb200::set_error("cuStreamWaitValue32 failed (%d)", static_cast<int>(r));
        return ctcb200::ERR_DRIVER;
    }
    return ctcb200::OK;
}

// Stream-ordered store to a device word (driver stream memory operation): *counter = value once everything enqueued on `stream`
// before this call has completed (and its writes are visible). The counterpart of ctcb200_stream_wait_geq for publishing the
// progress of a chunked producer to a kernel that is already running (ctcb200_lstm_fwd_streamed polls such a word).
extern "C" CTCB200_API int ctcb200_stream_write_value(ctcb200_stream_t stream_, void* counter, uint32_t value) {
    typedef CUresult (*WriteFn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
    static WriteFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuStreamWriteValue32", &p, cudaEnableDefault, &q);
        if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
            (void)cudaGetLastError();
            ctcb200::set_error("cuStreamWriteValue32 driver entry point unavailable");
            return ctcb200::ERR_DRIVER;
        }
        fn = reinterpret_cast<WriteFn>(p);
    }
    if (!counter || (reinterpret_cast<uintptr_t>(counter) & 3) != 0) {
        ctcb200::set_error("stream_write_value: counter %p must be a 4-byte aligned device address", counter);
        return ctcb200::ERR_INVALID;
    }
    CUresult r = fn(static_cast<CUstream>(stream_), reinterpret_cast<CUdeviceptr>(counter), value, 0u /* default: with memory barrier */);
    if (r != CUDA_SUCCESS) {
        ctcb200::set_error("cuStreamWriteValue32 failed (%d)", static_cast<int>(r));
        return ctcb200::ERR_DRIVER;
    }
    return ctcb200::OK;
}

// ---- can a kernel on one stream make progress while a kernel on another stream is resident? ----------------------------------
// The streamed input projection (ctcb200_lstm_fwd_streamed) launches a kernel that WAITS for the output of kernels launched after
// it on a second stream. That is only legal when the device really runs them concurrently: profilers (Nsight Compute serialises
// every kernel), CUDA_LAUNCH_BLOCKING, some debuggers and MPS configurations do not — the recurrent kernel would wait until its
// timeout trap. The probe runs the same pattern in miniature: a one-thread kernel on stream_a polls a flag for at most
// `limit_ms`; a trivial kernel followed by a stream memory operation on stream_b raises the flag.
namespace ctcb200 {
namespace {
__global__ void probe_wait_kernel(const unsigned int* flag, unsigned int* result, long long limit_cycles) {
    const long long t0 = clock64();
    while (ld_acquire_sys(flag) == 0u) {
        if (clock64() - t0 > limit_cycles) { *result = 0u; return; }
        __nanosleep(500);
    }
    *result = 1u;
}
__global__ void probe_touch_kernel(unsigned int* word) { *word = 1u; }
}  // namespace
}  // namespace ctcb200

// returns 1 (concurrent), 0 (the second stream's work only ran after the waiting kernel gave up) or a negative error code.
// Synchronises both streams; meant to be called once per device and process, not per step.
extern "C" CTCB200_API int ctcb200_concurrency_probe(ctcb200_stream_t stream_a_, ctcb200_stream_t stream_b_, int limit_ms) {
    using namespace ctcb200;
    cudaStream_t sa = static_cast<cudaStream_t>(stream_a_), sb = static_cast<cudaStream_t>(stream_b_);
    CTCB_REQUIRE(sa != sb, "concurrency_probe: the two streams must differ");
    CTCB_REQUIRE(limit_ms > 0 && limit_ms <= 1000, "concurrency_probe: limit_ms %d not in (0, 1000]", limit_ms);
    // both kernels loaded before the first one spins (a lazy first-launch load of the second would itself wait for the first)
    cudaFuncAttributes fa;
    CTCB_CUDA(cudaFuncGetAttributes(&fa, probe_wait_kernel));
    CTCB_CUDA(cudaFuncGetAttributes(&fa, probe_touch_kernel));
    int khz = 0;
    CTCB_CUDA(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, current_device()));
    if (khz <= 0) khz = 2000000;
    unsigned int* words = nullptr;   // [0] flag, [1] result, [2] scratch of the touch kernel
    CTCB_CUDA(cudaMalloc(&words, 3 * sizeof(unsigned int)));
    int rc = OK;
    unsigned int result = 0;
    cudaError_t e = cudaMemsetAsync(words, 0, 3 * sizeof(unsigned int), sa);
    if (e == cudaSuccess) e = cudaStreamSynchronize(sa);
    if (e == cudaSuccess) {
        probe_wait_kernel<<<1, 1, 0, sa>>>(words, words + 1, static_cast<long long>(khz) * limit_ms);
        probe_touch_kernel<<<1, 1, 0, sb>>>(words + 2);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) {
        rc = ctcb200_stream_write_value(sb, words, 1u);
        cudaError_t e1 = cudaStreamSynchronize(sa), e2 = cudaStreamSynchronize(sb);
        e = e1 != cudaSuccess ? e1 : e2;
    }
    if (e == cudaSuccess && rc == OK) e = cudaMemcpy(&result, words + 1, sizeof(result), cudaMemcpyDeviceToHost);
    (void)cudaFree(words);
    if (e != cudaSuccess) return cuda_fail(e, "concurrency probe", __FILE__, __LINE__);
    if (rc != OK) return rc;
    return result ? 1 : 0;
}
