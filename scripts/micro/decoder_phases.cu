// Microbenchmark for the round-2 plan: the synchronisation + data-movement SKELETON of a persistent
// decoder-step kernel, without the arithmetic.  One cooperative grid of one CTA per SM runs S steps; a step is four sequentially
// dependent phases handed off through L2 with release / acquire counters:
//   ATT  (all CTAs)      : stream this CTA's share of K and V (2 * B * T * P * 4 bytes in total, L2 resident) and write a context slice
//   C0   (first n0 CTAs) : wait for ATT of this step, fetch a (32 x K0) bf16 operand tile, write a (32 x 128) fp32 result tile
//   C1   (next n1 CTAs)  : wait for C0, fetch a (32 x K1) bf16 tile, write a (32 x 128) fp32 tile
//   Q    (same n1 CTAs)  : wait for C1 (the h1 exchange among the cell-1 CTAs), fetch (32 x DO) bf16, write (32 x 32) fp32
//   next ATT waits for Q.
// It answers, before the real kernel is written: how long does a step take when only the hand-offs and the memory traffic remain
// (the floor the design note estimates at ~15 us, against 31 us per forward step of the launch-per-stage loop), and how much of it is
// the ATT phase when persistent CTAs stream K/V with plain 16-byte loads.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o decoder_phases decoder_phases.cu
// Run:   ./decoder_phases [steps=300] [B=96] [T=200] [P=256] [n0=48] [n1=24] [att=1]      (att=0: skip the K/V streaming)
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

struct Args {
    const float4* kv;        // K then V: 2 * B * T * P floats
    long long kv_vec;        // float4 count
    const uint4* act0;       // cell-0 operand tiles: 3 batch slices x (32 x K0) bf16
    const uint4* act1;       // cell-1 operand tiles
    float* ctx;              // (B, P)
    float* out0;             // n0 x (32 x 128)
    float* out1;             // n1 x (32 x 128)
    float* outq;             // n1 x (32 x 32)
    unsigned* ctr;           // 4 counters, 128 bytes apart
    long long* stamps;       // optional: per-step clock of CTA 0 at the end of each phase (4 per step)
    int steps, n0, n1, K0, K1, DO, att;
};

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release(unsigned* p) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}
// all threads of the CTA have finished their global stores -> one release increment
__device__ __forceinline__ void cta_arrive(unsigned* ctr) {
    __syncthreads();
    if (threadIdx.x == 0) red_release(ctr);
}
// one thread polls, the CTA follows
__device__ __forceinline__ void cta_wait(const unsigned* ctr, unsigned target) {
    if (threadIdx.x == 0) {
        while ((int)(ld_acquire(ctr) - target) < 0) { }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256, 1) phases_kernel(const Args a) {
    extern __shared__ __align__(16) uint8_t sm[];
    uint4* tile = reinterpret_cast<uint4*>(sm);
    const int cta = blockIdx.x, ncta = gridDim.x, tid = threadIdx.x;
    unsigned* c_att = a.ctr;
    unsigned* c_c0 = a.ctr + 32;
    unsigned* c_c1 = a.ctr + 64;
    unsigned* c_q = a.ctr + 96;
    float acc = 0.f;
    for (int s = 0; s < a.steps; ++s) {
        // ---- ATT: every CTA, after the query of this step exists ----
        if (s > 0) cta_wait(c_q, (unsigned)(a.n1 * s));
        if (a.att) {
            const long long per = (a.kv_vec + ncta - 1) / ncta;
            const long long lo = per * cta, hi = (lo + per < a.kv_vec) ? lo + per : a.kv_vec;
            float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0, r2 = r0, r3 = r0;
            long long i = lo + tid;
            for (; i + 3 * 256 < hi; i += 4 * 256) {            // four independent 16-byte loads in flight per thread
                const float4 v0 = a.kv[i], v1 = a.kv[i + 256], v2 = a.kv[i + 512], v3 = a.kv[i + 768];
                r0.x += v0.x; r0.y += v0.y; r0.z += v0.z; r0.w += v0.w;
                r1.x += v1.x; r1.y += v1.y; r1.z += v1.z; r1.w += v1.w;
                r2.x += v2.x; r2.y += v2.y; r2.z += v2.z; r2.w += v2.w;
                r3.x += v3.x; r3.y += v3.y; r3.z += v3.z; r3.w += v3.w;
            }
            for (; i < hi; i += 256) { const float4 v = a.kv[i]; r0.x += v.x; r0.y += v.y; r0.z += v.z; r0.w += v.w; }
            acc += r0.x + r0.y + r0.z + r0.w + r1.x + r1.y + r1.z + r1.w + r2.x + r2.y + r2.z + r2.w + r3.x + r3.y + r3.z + r3.w;
        }
        a.ctx[(long long)cta * 256 + tid] = acc;                 // stands in for this CTA's context rows
        cta_arrive(c_att);
        if (cta == 0 && tid == 0 && a.stamps) a.stamps[4 * s + 0] = clock64();
        // ---- C0 ----
        if (cta < a.n0) {
            cta_wait(c_att, (unsigned)(ncta * (s + 1)));
            const int nvec = 32 * a.K0 * 2 / 16;
            const uint4* src = a.act0 + (long long)(cta % 3) * nvec;
            for (int i = tid; i < nvec; i += 256) tile[i] = src[i];
            __syncthreads();
            float v = 0.f;
            for (int i = tid; i < nvec; i += 256) v += __uint_as_float(tile[i].x & 0xffff0000u);
            for (int i = tid; i < 32 * 128; i += 256) a.out0[(long long)cta * 32 * 128 + i] = v + acc;
            cta_arrive(c_c0);
            if (cta == 0 && tid == 0 && a.stamps) a.stamps[4 * s + 1] = clock64();
        }
        // ---- C1 and Q ----
        if (cta >= a.n0 && cta < a.n0 + a.n1) {
            const int c1 = cta - a.n0;
            cta_wait(c_c0, (unsigned)(a.n0 * (s + 1)));
            int nvec = 32 * a.K1 * 2 / 16;
            const uint4* src = a.act1 + (long long)(c1 % 3) * nvec;
            for (int i = tid; i < nvec; i += 256) tile[i] = src[i];
            __syncthreads();
            float v = 0.f;
            for (int i = tid; i < nvec; i += 256) v += __uint_as_float(tile[i].x & 0xffff0000u);
            for (int i = tid; i < 32 * 128; i += 256) a.out1[(long long)c1 * 32 * 128 + i] = v + acc;
            cta_arrive(c_c1);
            cta_wait(c_c1, (unsigned)(a.n1 * (s + 1)));
            nvec = 32 * a.DO * 2 / 16;
            for (int i = tid; i < nvec; i += 256) tile[i] = reinterpret_cast<const uint4*>(a.out1)[(long long)(c1 % 3) * nvec + i];
            __syncthreads();
            v = 0.f;
            for (int i = tid; i < nvec; i += 256) v += __uint_as_float(tile[i].x);
            for (int i = tid; i < 32 * 32; i += 256) a.outq[(long long)c1 * 32 * 32 + i] = v;
            cta_arrive(c_q);
        }
        if (cta == 0 && tid == 0 && a.stamps) a.stamps[4 * s + 2] = clock64();
    }
}

int main(int argc, char** argv) {
    const int steps = argc > 1 ? atoi(argv[1]) : 300;
    const int B = argc > 2 ? atoi(argv[2]) : 96, T = argc > 3 ? atoi(argv[3]) : 200, P = argc > 4 ? atoi(argv[4]) : 256;
    const int n0 = argc > 5 ? atoi(argv[5]) : 48, n1 = argc > 6 ? atoi(argv[6]) : 24, att = argc > 7 ? atoi(argv[7]) : 1;
    const int K0 = 768, K1 = 768, DO = 256;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int ncta = prop.multiProcessorCount;
    if (n0 + n1 > ncta) { printf("n0 + n1 must be <= %d\n", ncta); return 1; }
    Args a{};
    const long long kv_floats = 2LL * B * T * P;
    float* kv; CK(cudaMalloc(&kv, kv_floats * 4)); CK(cudaMemset(kv, 0, kv_floats * 4));
    uint4 *act0, *act1; CK(cudaMalloc(&act0, 3LL * 32 * K0 * 2)); CK(cudaMalloc(&act1, 3LL * 32 * K1 * 2));
    CK(cudaMemset(act0, 0, 3LL * 32 * K0 * 2)); CK(cudaMemset(act1, 0, 3LL * 32 * K1 * 2));
    float *ctx, *out0, *out1, *outq;
    CK(cudaMalloc(&ctx, (size_t)ncta * 256 * 4)); CK(cudaMalloc(&out0, (size_t)n0 * 32 * 128 * 4));
    CK(cudaMalloc(&out1, (size_t)(n1 > 3 ? n1 : 3) * 32 * 128 * 4)); CK(cudaMalloc(&outq, (size_t)n1 * 32 * 32 * 4));
    CK(cudaMemset(out1, 0, (size_t)(n1 > 3 ? n1 : 3) * 32 * 128 * 4));
    unsigned* ctr; CK(cudaMalloc(&ctr, 512));
    a.kv = (const float4*)kv; a.kv_vec = kv_floats / 4; a.act0 = act0; a.act1 = act1; a.ctx = ctx; a.out0 = out0; a.out1 = out1; a.outq = outq;
    a.ctr = ctr; a.stamps = nullptr; a.steps = steps; a.n0 = n0; a.n1 = n1; a.K0 = K0; a.K1 = K1; a.DO = DO; a.att = att;
    const size_t smem = (size_t)32 * K0 * 2;
    CK(cudaFuncSetAttribute(phases_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, phases_kernel, 256, smem));
    if (per_sm < 1) { printf("kernel does not fit\n"); return 1; }
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    void* args[] = {(void*)&a};
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaMemset(ctr, 0, 512));
        CK(cudaEventRecord(e0));
        CK(cudaLaunchCooperativeKernel((void*)phases_kernel, dim3(ncta), dim3(256), args, smem, 0));
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    const double us = best * 1e3 / steps;
    printf("decoder phase skeleton: %d CTAs, %d steps, B=%d T=%d P=%d (K+V %.1f MB), cell-0 CTAs %d, cell-1 CTAs %d, att=%d\n", ncta, steps, B, T, P,
           kv_floats * 4 / 1e6, n0, n1, att);
    printf("  %.2f us per step", us);
    if (att) printf("  (K/V streaming alone would be %.2f us at %.0f GB/s)", kv_floats * 4 / 1e3 / 4800.0, 4800.0);
    printf("\n");
    return 0;
}
