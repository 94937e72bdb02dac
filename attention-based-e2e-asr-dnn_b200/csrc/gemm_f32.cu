// fp32 SIMT GEMM with two-level strided indexing (parity mode: 1e-4 relative error against the reference needs
// true fp32 accumulation, so this path uses FFMA, not tensor cores).
//
//   C[m][n] = alpha * sum_k A(m,k) * B(k,n) + beta * C[m][n] + bias1[n] + bias2[n]
//
// The two-level index (i -> (i / inner) * s_outer + (i % inner) * s_inner) is what lets the pyramidal
// frame-pair concat (reference src/modules.py:171-185), the odd-frame truncation and the "output length is
// max(lx)" rule be pure addressing instead of copies.
#include "las_common.cuh"
#include "las_b200.h"
#include <stdint.h>

namespace {

struct Idx2 {
    long long so, si;
    int inner;   // 0 => single level
    __device__ __forceinline__ long long operator()(int i) const {
        if (inner == 0) return (long long)i * si;
        int q = i / inner;
        return (long long)q * so + (long long)(i - q * inner) * si;
    }
};

struct GemmArgs {
    const float* A;
    const float* B;
    float* C;
    const float* bias1;
    const float* bias2;
    int M, N, K;
    Idx2 am, ak, bk, cm;
    long long bn;
    long long bsA, bsB, bsC;
    float alpha, beta;
    int a_kfast, b_nfast;
    int c_vec4;          // every C row start, the batch stride and N are multiples of 4 floats and C is 16-byte aligned: 128-bit stores
};

template <int BM, int BN, int BK, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN), ((BM / TM) * (BN / TN) >= 256) ? 2 : ((BM / TM) * (BN / TN) == 128 ? 3 : 1)) gemm_f32_kernel(GemmArgs g) {
    // 256-thread tiles: two CTAs per SM (128 registers) -- one CTA is 2 warps per scheduler, too few to cover the shared-memory latency
    constexpr int NT = (BM / TM) * (BN / TN);
    constexpr int PAD = 4;
    __shared__ __align__(16) float As[2][BK][BM + PAD];
    __shared__ __align__(16) float Bs[2][BK][BN + PAD];
    constexpr int A_PER = (BM * BK) / NT;
    constexpr int B_PER = (BN * BK) / NT;
    static_assert((BM * BK) % NT == 0 && (BN * BK) % NT == 0, "tile/threads mismatch");
    static_assert(TM % 4 == 0 && TN % 4 == 0, "micro tile must be float4-able");

    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const float* A = g.A + (long long)blockIdx.z * g.bsA;
    const float* B = g.B + (long long)blockIdx.z * g.bsB;
    float* C = g.C + (long long)blockIdx.z * g.bsC;

    // loader coordinates
    int a_m[A_PER], a_k[A_PER], b_n[B_PER], b_k[B_PER];
    long long a_off[A_PER], b_off[B_PER];
#pragma unroll
    for (int i = 0; i < A_PER; ++i) {
        int e = tid + i * NT;
        if (g.a_kfast) { a_k[i] = e % BK; a_m[i] = e / BK; }
        else           { a_m[i] = e % BM; a_k[i] = e / BM; }
        int m = m0 + a_m[i];
        a_off[i] = (m < g.M) ? g.am(m) : -1;
    }
#pragma unroll
    for (int i = 0; i < B_PER; ++i) {
        int e = tid + i * NT;
        if (g.b_nfast) { b_n[i] = e % BN; b_k[i] = e / BN; }
        else           { b_k[i] = e % BK; b_n[i] = e / BK; }
        int n = n0 + b_n[i];
        b_off[i] = (n < g.N) ? (long long)n * g.bn : -1;
    }

    float ra[A_PER], rb[B_PER];
    auto load_tile = [&](int k0) {
#pragma unroll
        for (int i = 0; i < A_PER; ++i) {
            int k = k0 + a_k[i];
            ra[i] = (a_off[i] >= 0 && k < g.K) ? __ldg(A + a_off[i] + g.ak(k)) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < B_PER; ++i) {
            int k = k0 + b_k[i];
            rb[i] = (b_off[i] >= 0 && k < g.K) ? __ldg(B + b_off[i] + g.bk(k)) : 0.f;
        }
    };
    auto store_tile = [&](int buf) {
#pragma unroll
        for (int i = 0; i < A_PER; ++i) As[buf][a_k[i]][a_m[i]] = ra[i];
#pragma unroll
        for (int i = 0; i < B_PER; ++i) Bs[buf][b_k[i]][b_n[i]] = rb[i];
    };

    // compute mapping: thread owns TM rows as TM/4 groups of 4 (group j at ty*4 + j*(BM*4/TM)), same for cols
    constexpr int TX = BN / TN, TY = BM / TM;
    const int tx = tid % TX, ty = tid / TX;
    constexpr int GM = TM / 4, GN = TN / 4;
    constexpr int SM_ = BM / GM, SN_ = BN / GN;   // spacing between a thread's groups
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    const int nk = ceil_div(g.K, BK);
    load_tile(0);
    store_tile(0);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) load_tile((kt + 1) * BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float av[TM], bv[TN];
#pragma unroll
            for (int j = 0; j < GM; ++j) {
                float4 v = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4 + j * SM_]);
                av[j * 4 + 0] = v.x; av[j * 4 + 1] = v.y; av[j * 4 + 2] = v.z; av[j * 4 + 3] = v.w;
            }
#pragma unroll
            for (int j = 0; j < GN; ++j) {
                float4 v = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4 + j * SN_]);
                bv[j * 4 + 0] = v.x; bv[j * 4 + 1] = v.y; bv[j * 4 + 2] = v.z; bv[j * 4 + 3] = v.w;
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (kt + 1 < nk) {
            store_tile(buf ^ 1);
            __syncthreads();
        }
    }

    // epilogue
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int m = m0 + ty * 4 + (i / 4) * SM_ + (i % 4);
        if (m >= g.M) continue;
        float* crow = C + g.cm(m);
        if (g.c_vec4) {
            // a thread's columns come in groups of four consecutive ones: one 128-bit store per group instead of four scalar stores
            // whose lanes sit 16 bytes apart (a (28800 x 512) output with K = 30 -- the decoder's dQC -- is all epilogue)
#pragma unroll
            for (int jg = 0; jg < GN; ++jg) {
                const int n = n0 + tx * 4 + jg * SN_;
                if (n >= g.N) continue;          // N % 4 == 0: the group is entirely inside or entirely outside
                float4 v = make_float4(g.alpha * acc[i][jg * 4 + 0], g.alpha * acc[i][jg * 4 + 1], g.alpha * acc[i][jg * 4 + 2],
                                       g.alpha * acc[i][jg * 4 + 3]);
                if (g.bias1) { v.x += __ldg(g.bias1 + n); v.y += __ldg(g.bias1 + n + 1); v.z += __ldg(g.bias1 + n + 2); v.w += __ldg(g.bias1 + n + 3); }
                if (g.bias2) { v.x += __ldg(g.bias2 + n); v.y += __ldg(g.bias2 + n + 1); v.z += __ldg(g.bias2 + n + 2); v.w += __ldg(g.bias2 + n + 3); }
                if (g.beta != 0.f) {
                    const float4 o = *reinterpret_cast<const float4*>(crow + n);
                    v.x += g.beta * o.x; v.y += g.beta * o.y; v.z += g.beta * o.z; v.w += g.beta * o.w;
                }
                *reinterpret_cast<float4*>(crow + n) = v;
            }
            continue;
        }
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            int n = n0 + tx * 4 + (j / 4) * SN_ + (j % 4);
            if (n >= g.N) continue;
            float v = g.alpha * acc[i][j];
            if (g.bias1) v += __ldg(g.bias1 + n);
            if (g.bias2) v += __ldg(g.bias2 + n);
            if (g.beta != 0.f) v += g.beta * crow[n];
            crow[n] = v;
        }
    }
}

template <int BM, int BN, int BK, int TM, int TN>
int launch(const GemmArgs& g, int batch, cudaStream_t st) {
    dim3 grid(ceil_div(g.N, BN), ceil_div(g.M, BM), batch);
    dim3 block((BM / TM) * (BN / TN));
    gemm_f32_kernel<BM, BN, BK, TM, TN><<<grid, block, 0, st>>>(g);
    LAS_LAUNCH_CHECK();
    return LAS_OK;
}

}  // namespace

extern "C" int las_gemm_f32(const LasGemmF32* d, void* stream) {
    LAS_CHECK_ARG(d != nullptr, "las_gemm_f32: null descriptor");
    LAS_CHECK_ARG(d->M >= 0 && d->N >= 0 && d->K >= 0 && d->batch >= 1, "las_gemm_f32: bad dims M=%d N=%d K=%d batch=%d",
                  d->M, d->N, d->K, d->batch);
    if (d->M == 0 || d->N == 0) return LAS_OK;
    LAS_CHECK_ARG(d->A && d->B && d->C, "las_gemm_f32: null operand");
    LAS_CHECK_ARG(d->batch <= 65535, "las_gemm_f32: batch %d too large", d->batch);
    int rc = las_set_device_of(d->C);
    if (rc) return rc;
    GemmArgs g;
    g.A = d->A; g.B = d->B; g.C = d->C; g.bias1 = d->bias1; g.bias2 = d->bias2;
    g.M = d->M; g.N = d->N; g.K = d->K;
    g.am = {d->a_m_so, d->a_m_si, d->a_m_inner};
    g.ak = {d->a_k_so, d->a_k_si, d->a_k_inner};
    g.bk = {d->b_k_so, d->b_k_si, d->b_k_inner};
    g.cm = {d->c_m_so, d->c_m_si, d->c_m_inner};
    g.bn = d->b_n_s;
    g.bsA = d->bsA; g.bsB = d->bsB; g.bsC = d->bsC;
    g.alpha = d->alpha; g.beta = d->beta;
    g.a_kfast = (d->a_k_si == 1) ? 1 : 0;
    g.b_nfast = (d->b_n_s == 1) ? 1 : 0;
    g.c_vec4 = (d->N % 4 == 0 && d->c_m_si % 4 == 0 && (d->c_m_inner == 0 || d->c_m_so % 4 == 0) && d->bsC % 4 == 0 &&
                ((uintptr_t)d->C & 15) == 0) ? 1 : 0;
    cudaStream_t st = (cudaStream_t)stream;
    LasProfScope prof(d->prof_tag == 1 ? LAS_PROF_GEMM_GATES : LAS_PROF_GEMM_OTHER, stream,
                      2.0 * d->M * (double)d->N * d->K * d->batch);
    const LasDeviceInfo* di = las_device_info();
    // a tall output with at most 32 columns (the tied classifier over all decoder steps: 28800 x 30, K = 512) would spend three quarters
    // of a 128-column tile on padding
    if (g.N <= 32 && (long long)ceil_div(g.M, 128) * d->batch >= di->num_sms) return launch<128, 32, 16, 8, 4>(g, d->batch, st);
    long long big_tiles = (long long)ceil_div(g.M, 128) * ceil_div(g.N, 128) * d->batch;
    if (big_tiles >= di->num_sms) return launch<128, 128, 8, 8, 8>(g, d->batch, st);
    long long mid_tiles = (long long)ceil_div(g.M, 64) * ceil_div(g.N, 64) * d->batch;
    if (mid_tiles >= di->num_sms / 2) return launch<64, 64, 16, 4, 4>(g, d->batch, st);
    return launch<32, 32, 32, 4, 4>(g, d->batch, st);
}
