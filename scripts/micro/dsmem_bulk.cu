// Microbenchmark: all-to-all exchange of 2 KB slices inside a 16-CTA thread-block cluster with bulk async DSMEM copies
// (cp.async.bulk.shared::cluster.shared::cta + remote mbarrier complete_tx) -- the h_t exchange pattern of the LSTM recurrence.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dsmem_bulk dsmem_bulk.cu ; run: ./dsmem_bulk [cluster=16] [slice_bytes=2048]
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do { asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory"); } while (!done);
}
__device__ __forceinline__ void bulk_s2s(uint32_t dst_cluster_addr, uint32_t src_cta_addr, uint32_t bytes, uint32_t remote_bar) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_cluster_addr), "r"(src_cta_addr), "r"(bytes), "r"(remote_bar) : "memory");
}

__global__ void __launch_bounds__(128, 1) exch(int steps, int slice, long long* out, unsigned* check) {
    extern __shared__ __align__(1024) uint8_t sm[];
    uint32_t rank, csz;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(csz));
    // layout: stage[slice] | recv[2][csz*slice] | bars[2]
    uint8_t* stage = sm;
    uint8_t* recv = sm + slice;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + slice + 2 * csz * slice);
    const uint32_t bar0 = smem_u32(bars), recv0 = smem_u32(recv), stage0 = smem_u32(stage);
    if (threadIdx.x == 0) { mbar_init(bar0, 1); mbar_init(bar0 + 8, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    long long t_issue = 0, t_total = 0;
    unsigned acc = 0;
    for (int s = 0; s < steps; ++s) {
        const int par = s & 1;
        // "compute": fill the stage tile
        for (int i = threadIdx.x; i < slice / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(stage)[i] = rank * 1000003u + s * 17u + i;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        long long t0 = clock64();
        if (threadIdx.x == 0) {
            mbar_expect_tx(bar0 + 8 * par, csz * slice);                     // I will receive csz slices this step
            for (uint32_t d = 0; d < csz; ++d) {
                const uint32_t dst = mapa(recv0 + par * csz * slice + rank * slice, d);
                const uint32_t rbar = mapa(bar0 + 8 * par, d);
                bulk_s2s(dst, stage0, slice, rbar);
            }
        }
        long long t1 = clock64();
        mbar_wait(bar0 + 8 * par, (s >> 1) & 1);
        long long t2 = clock64();
        // consume: checksum one word from every source
        for (uint32_t r2 = threadIdx.x; r2 < csz; r2 += blockDim.x) acc += reinterpret_cast<uint32_t*>(recv + par * csz * slice + r2 * slice)[1] - (r2 * 1000003u + s * 17u + 1);
        if (threadIdx.x == 0) { asm volatile("cp.async.bulk.commit_group;\ncp.async.bulk.wait_group.read 0;" ::: "memory"); }
        __syncthreads();
        if (s >= 8) { t_issue += t1 - t0; t_total += t2 - t0; }
    }
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (threadIdx.x == 0) { out[blockIdx.x * 2] = t_issue / (steps - 8); out[blockIdx.x * 2 + 1] = t_total / (steps - 8); }
    atomicAdd(check, acc);
}

int main(int argc, char** argv) {
    int csz = argc > 1 ? atoi(argv[1]) : 16, slice = argc > 2 ? atoi(argv[2]) : 2048, nclusters = argc > 3 ? atoi(argv[3]) : 6, steps = 200;
    size_t smem = slice + 2 * (size_t)csz * slice + 64;
    cudaFuncSetAttribute(exch, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (csz > 8) cudaFuncSetAttribute(exch, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    long long* out; unsigned* chk;
    cudaMalloc(&out, sizeof(long long) * 2 * csz * nclusters); cudaMalloc(&chk, 4); cudaMemset(chk, 0, 4);
    cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(csz * nclusters); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = csz; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int maxc = 0; cudaOccupancyMaxActiveClusters(&maxc, exch, &cfg);
    cudaError_t e = cudaLaunchKernelEx(&cfg, exch, steps, slice, out, chk);
    cudaError_t e2 = cudaDeviceSynchronize();
    long long h[2 * 16 * 16]; unsigned hc = 1;
    cudaMemcpy(h, out, sizeof(long long) * 2 * csz * nclusters, cudaMemcpyDeviceToHost); cudaMemcpy(&hc, chk, 4, cudaMemcpyDeviceToHost);
    printf("cluster %d slice %d B clusters %d (max active %d): launch %s sync %s checksum_err %u\n", csz, slice, nclusters, maxc, cudaGetErrorString(e), cudaGetErrorString(e2), hc);
    long long mi = 1LL << 60, ma = 0, mi2 = 1LL << 60, ma2 = 0;
    for (int i = 0; i < csz * nclusters; ++i) { if (h[2*i] < mi) mi = h[2*i]; if (h[2*i] > ma) ma = h[2*i]; if (h[2*i+1] < mi2) mi2 = h[2*i+1]; if (h[2*i+1] > ma2) ma2 = h[2*i+1]; }
    printf("  issue cycles/step: min %lld max %lld ; issue->all received: min %lld max %lld  (%.1f B/clk ingest per SM)\n", mi, ma, mi2, ma2, (double)csz * slice / ma2);
    return 0;
}
