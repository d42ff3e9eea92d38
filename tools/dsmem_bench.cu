// DSMEM (SM-to-SM shared memory) microbenchmark for the exchange pattern of the recurrent kernels (csrc/lstm.cu):
// every CTA of a thread-block cluster bulk-copies one block of `bytes` to EVERY CTA of the cluster
// (cp.async.bulk.shared::cluster.shared::cta with complete_tx on the receiver's mbarrier = an all-gather), waits until all
// blocks addressed to it have landed, and repeats. Reports cycles per exchange and bytes/clk per SM (in + out) so that the
// "DSMEM-bound" claim of DESIGN.md §3.2 rests on a number measured on THIS chip.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dsmem_bench.bin tools/dsmem_bench.cu
//   tools/dsmem_bench.bin            -> JSON lines on stdout (profiles/dsmem_microbench.json keeps the summary)
//
// Modes: "allgather" (bandwidth: all copies in flight, one wait per exchange) and "pingpong" (latency: CTA 0 -> CTA 1 -> CTA 0
// with one small copy each way: the hand-off floor of one recurrent time step).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s failed: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
    return r;
}
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory");
}
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void bulk_copy(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t bar_cluster) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_cluster), "r"(src_cta), "r"(bytes), "r"(bar_cluster) : "memory");
}

// all-gather: per exchange every CTA sends `bytes` to each of the `cs` CTAs (itself included, like the recurrent kernels)
__global__ void __launch_bounds__(256, 1) allgather_kernel(int bytes, int iters, long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t cs = gridDim.x;                 // one cluster = the whole x extent
    uint8_t* recv = smem;                          // [2 parities][cs][bytes]
    uint8_t* send = smem + 2 * cs * bytes;         // [bytes]
    uint64_t* bar = reinterpret_cast<uint64_t*>(send + bytes);   // [2]
    const uint32_t me = cluster_rank();
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect(&bar[0], cs * bytes);
        mbar_expect(&bar[1], cs * bytes);
    }
    for (int i = tid; i < bytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(send)[i] = me * 1000 + i;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    cluster_sync();
    long long t0 = 0;
    for (int it = 0; it < iters; ++it) {
        if (it == 8 && tid == 0) t0 = clock64();
        const int par = it & 1;
        // warp w issues the copies to peers w, w + 8 (as the recurrent kernels do)
        if (lane == 0)
            for (uint32_t d = warp; d < cs; d += 8)
                bulk_copy(mapa(smem_u32(recv + (par * cs + me) * bytes), d), smem_u32(send), bytes, mapa(smem_u32(&bar[par]), d));
        // everybody waits for the cs blocks addressed to this CTA, then one thread re-arms the barrier for exchange it + 2
        mbar_wait(&bar[par], (it >> 1) & 1);
        __syncthreads();
        if (tid == 0) mbar_expect(&bar[par], cs * bytes);
        // the next exchange into this parity happens two iterations later: by then every peer has passed the wait of the
        // exchange in between, which needed our copies of that exchange, issued after this point: no overwrite hazard
    }
    if (tid == 0 && blockIdx.x == 0 && blockIdx.y == 0) cycles[0] = clock64() - t0;
    __syncthreads();
    cluster_sync();
}


// all-gather with per-lane remote stores instead of the bulk-copy engine: MODE 1 = st.async (16 B per lane, complete_tx on the
// receiver's mbarrier), MODE 2 = st.shared::cluster.v4 followed by one release-arrive per (sender warp, peer) on the receiver's
// mbarrier. `send_warps` warps share the peers round-robin. The receiver fences the async proxy after the wait (what a
// tcgen05.mma consumer of the image would need).
__device__ __forceinline__ void st_async_v4(uint32_t dst_cluster, uint4 v, uint32_t bar_cluster) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                 ::"r"(dst_cluster), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void st_cluster_v4(uint32_t dst_cluster, uint4 v) {
    asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst_cluster), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void remote_arrive_release(uint32_t bar_cluster) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}
template <int MODE>
__global__ void __launch_bounds__(256, 1) allgather_st_kernel(int bytes, int iters, int send_warps, long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t cs = gridDim.x;
    uint8_t* recv = smem;                          // [2 parities][cs][bytes]
    uint8_t* send = smem + 2 * cs * bytes;         // [bytes]
    uint64_t* bar = reinterpret_cast<uint64_t*>(send + bytes);   // [2]
    const uint32_t me = cluster_rank();
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        mbar_init(&bar[0], MODE == 1 ? 1 : cs);
        mbar_init(&bar[1], MODE == 1 ? 1 : cs);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (MODE == 1) { mbar_expect(&bar[0], cs * bytes); mbar_expect(&bar[1], cs * bytes); }
    }
    for (int i = tid; i < bytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(send)[i] = me * 1000 + i;
    __syncthreads();
    cluster_sync();
    long long t0 = 0;
    uint32_t check = 0;
    for (int it = 0; it < iters; ++it) {
        if (it == 8 && tid == 0) t0 = clock64();
        const int par = it & 1;
        if (warp < send_warps) {
            for (uint32_t d = warp; d < cs; d += send_warps) {
                const uint32_t dst = mapa(smem_u32(recv + (par * cs + me) * bytes), d);
                const uint32_t rb = mapa(smem_u32(&bar[par]), d);
                for (int o = lane * 16; o < bytes; o += 512) {
                    const uint4 v = *reinterpret_cast<const uint4*>(send + o);
                    if (MODE == 1) st_async_v4(dst + o, v, rb); else st_cluster_v4(dst + o, v);
                }
                if (MODE == 2) { __syncwarp(); if (lane == 0) remote_arrive_release(rb); }
            }
        }
        mbar_wait(&bar[par], (it >> 1) & 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        check += reinterpret_cast<const uint32_t*>(recv + (par * cs + (tid % cs)) * bytes)[0];
        __syncthreads();
        if (MODE == 1 && tid == 0) mbar_expect(&bar[par], cs * bytes);
    }
    if (tid == 0 && blockIdx.x == 0 && blockIdx.y == 0) cycles[0] = clock64() - t0;
    if (check == 0xffffffffu) cycles[1] = check;
    __syncthreads();
    cluster_sync();
}

// ping-pong between CTA 0 and CTA 1 of a cluster: the latency of one bulk-copy hand-off
__global__ void __launch_bounds__(32, 1) pingpong_kernel(int bytes, int iters, long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* recv = smem;
    uint8_t* send = smem + bytes;
    uint64_t* bar = reinterpret_cast<uint64_t*>(send + bytes);
    const uint32_t me = cluster_rank();
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect(bar, bytes);
    }
    for (int i = threadIdx.x; i < bytes / 4; i += 32) reinterpret_cast<uint32_t*>(send)[i] = i;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    cluster_sync();
    long long t0 = clock64();
    if (me < 2 && threadIdx.x == 0) {
        const uint32_t peer = 1 - me;
        for (int it = 0; it < iters; ++it) {
            if (me == 0) {
                bulk_copy(mapa(smem_u32(recv), peer), smem_u32(send), bytes, mapa(smem_u32(bar), peer));
                mbar_wait(bar, it & 1);
                mbar_expect(bar, bytes);
            } else {
                mbar_wait(bar, it & 1);
                mbar_expect(bar, bytes);
                bulk_copy(mapa(smem_u32(recv), peer), smem_u32(send), bytes, mapa(smem_u32(bar), peer));
            }
        }
    }
    if (me == 0 && threadIdx.x == 0 && blockIdx.y == 0) cycles[0] = clock64() - t0;
    __syncwarp();
    cluster_sync();
}

template <typename K, typename... Extra>
static bool launch(K kern, int cs, int clusters, int threads, size_t smem, long long* d_cycles, int bytes, int iters, Extra... extra) {
    CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    if (cs > 8) CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cs, clusters, 1);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = cs; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n < clusters) {
        (void)cudaGetLastError();
        return false;
    }
    CHECK(cudaLaunchKernelEx(&cfg, kern, bytes, iters, extra..., d_cycles));
    CHECK(cudaDeviceSynchronize());
    return true;
}

int main() {
    cudaDeviceProp prop;
    CHECK(cudaGetDeviceProperties(&prop, 0));
    int clk_khz = 0;
    CHECK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    long long* d_cycles;
    CHECK(cudaMalloc(&d_cycles, 2 * sizeof(long long)));
    const int iters = 4008;
    printf("{\"device\": \"%s\", \"sms\": %d, \"sm_clock_khz\": %d}\n", prop.name, prop.multiProcessorCount, clk_khz);
    const int css[] = {16, 8, 4};
    const int sizes[] = {256, 512, 1024, 2048, 4096};
    for (int cs : css) {
        for (int clusters : {1, 4, 8}) {
            if (cs * clusters > prop.multiProcessorCount) continue;
            for (int bytes : sizes) {
                const size_t smem = static_cast<size_t>(2) * cs * bytes + bytes + 64 + 1024;
                if (smem > 200 * 1024) continue;
                if (!launch(allgather_kernel, cs, clusters, 256, smem, d_cycles, bytes, iters)) continue;
                long long cyc = 0;
                CHECK(cudaMemcpy(&cyc, d_cycles, sizeof(cyc), cudaMemcpyDeviceToHost));
                const double per = static_cast<double>(cyc) / (iters - 8);
                const double out_b = static_cast<double>(cs) * bytes;   // sent per SM per exchange (= received per SM)
                printf("{\"mode\": \"allgather\", \"cluster\": %d, \"clusters_resident\": %d, \"bytes_per_peer\": %d, "
                       "\"bytes_out_per_sm\": %.0f, \"cycles_per_exchange\": %.1f, \"B_per_clk_per_sm_in_plus_out\": %.2f}\n",
                       cs, clusters, bytes, out_b, per, 2.0 * out_b / per);
            }
        }
    }

    // the same all-gather with per-lane remote stores (no bulk-copy engine)
    for (int mode : {1, 2}) {
        for (int cs : {16, 8}) {
            for (int clusters : {4, 8}) {
                if (cs * clusters > prop.multiProcessorCount) continue;
                for (int bytes : {256, 512, 1024}) {
                    for (int sw : {1, 2, 4, 8}) {
                        const size_t smem = static_cast<size_t>(2) * cs * bytes + bytes + 64 + 1024;
                        const bool ok = mode == 1 ? launch(allgather_st_kernel<1>, cs, clusters, 256, smem, d_cycles, bytes, iters, sw)
                                                  : launch(allgather_st_kernel<2>, cs, clusters, 256, smem, d_cycles, bytes, iters, sw);
                        if (!ok) continue;
                        long long cyc = 0;
                        CHECK(cudaMemcpy(&cyc, d_cycles, sizeof(cyc), cudaMemcpyDeviceToHost));
                        const double per = static_cast<double>(cyc) / (iters - 8);
                        printf("{\"mode\": \"%s\", \"cluster\": %d, \"clusters_resident\": %d, \"bytes_per_peer\": %d, \"send_warps\": %d, "
                               "\"cycles_per_exchange\": %.1f}\n", mode == 1 ? "allgather_st_async" : "allgather_st_arrive", cs, clusters,
                               bytes, sw, per);
                    }
                }
            }
        }
    }
    for (int bytes : {16, 512, 1024}) {
        const size_t smem = static_cast<size_t>(2) * bytes + 64 + 1024;
        if (!launch(pingpong_kernel, 2, 1, 32, smem, d_cycles, bytes, 2000)) continue;
        long long cyc = 0;
        CHECK(cudaMemcpy(&cyc, d_cycles, sizeof(cyc), cudaMemcpyDeviceToHost));
        printf("{\"mode\": \"pingpong\", \"bytes\": %d, \"cycles_one_way\": %.1f}\n", bytes, static_cast<double>(cyc) / 2000 / 2);
    }
    return 0;
}
