// Speller decoder loop (forward + backward), fp32 parity mode.
//
// Replaces the Python time loop of Speller.forward (reference src/models.py:336-385): embedding lookup / teacher
// forcing select (:354-358), AutoRegDecoderLSTMCell.forward (src/modules.py:340-365: cat[emb, ctx] -> LSTMCell ->
// Dropout (dropped h is the recurrent state) -> LSTMCell -> Dropout), query_map + attention (:366), the tied
// classifier on cat[q_proj, ctx] (:370-373) and the greedy argmax feedback (:380).  The host enqueues every kernel of
// every step on one stream without ever synchronising; per-step history is laid out so that all weight gradients are
// single GEMMs over (steps*B) rows after the backward loop.
//
// Restructuring that keeps the arithmetic equivalent:
//   * cell-0 input GEMM: [emb, ctx, h0] . [W_ih0 | W_hh0]^T is split into a per-forward table
//     Gemb = emb . W_ih0[:, :E]^T + b_ih0 + b_hh0  (V x 4DH; the embedding has only V=30 rows) that is gathered by token,
//     plus one GEMM over the packed row S0[t] = [ctx_t | h0_{t-1}] with Wcat0 = [W_ih0[:, E:] | W_hh0];
//   * cell-1: packed row S1[t] = [h0_t | h1_{t-1}], Wcat1 = [W_ih1 | W_hh1];
//   * QC[t] = [q_proj_t | ctx_t] is both the classifier input and where attention reads its query.
#include "las_common.cuh"
#include "las_b200.h"
#include "dec_persist.h"
#include "attn_tail.h"
#include <vector>
#include <mutex>
#include <memory>
#include <stdlib.h>
#include <string.h>

// prepared tensor-core GEMM plans (gemm_tc.cu): tensor maps encoded once per loop, one launch per step
size_t las_tc_plan_bytes();
int las_tc_plan_make(void* plan_mem, const void* A, const void* B, int M, int N, int K, int a_batches, long long a_s1, long long a_s2,
                     long long b_s1, int b_mn_major);
int las_tc_plan_launch(const void* plan_mem, int a_batch, float* C, long long ldc, const float* bias1, const float* bias2, void* stream);
int las_tc_plan_launch_lstm(const void* plan_mem, int a_batch, const LasLstmEpi* le, void* stream);
int las_tc_plan_launch_split(const void* plan_mem, int a_batch, float* Cpart, long long ldp, int splitk, void* stream);
int las_tc_plan_tiles(const void* plan_mem);
int las_tc_plan_kiters(const void* plan_mem);
int las_permute_cast_lstm_rows(const float* src, long long ld_src, void* dst, int H, int K, void* stream);
int las_transpose_cast_f16(const float* src, void* dst, int rows, int cols, void* stream);   // decoder_persist.cu

namespace {

struct alignas(64) PlanBuf { unsigned char b[1024]; };
constexpr int MAX_SPLIT = 8;          // most split-K partials a per-step GEMM leaves for its pointwise consumer to add up

// ------------------------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------------------------
struct CellFwd {
    const float* Gin;         // optional: nsplit un-reduced split-K partials (nsplit, B, 4H), split_stride floats apart, added up here
    int nsplit; long long split_stride;
    const float* bias1; const float* bias2;   // optional (4H) biases (a split-K GEMM cannot add them)
    float* G;                 // (B, 4H) in: partial pre-activations (when Gin is null) ; out: activated gates
    const float* Gtab;        // (V, 4H) or null: row gathered by token and added
    const int* y; long long ld_y;   // gold tokens (B, >=steps) or null
    const int* chars_prev;    // (B) argmax of the previous step or null
    int* tok_out;             // (B) token actually fed (saved for backward) or null
    int t, use_gold, sos_idx;
    const float* c_prev; long long ld_cp;
    float* c_out; long long ld_co;
    const float* mask;        // (B, H) or null
    float* h1; long long ld_h1;
    float* h2; long long ld_h2;   // nullable second destination
    __nv_bfloat16* h1b; long long ld_h1b;   // nullable bf16 copies (tensor-pipe mode: next GEMMs' A operand)
    __nv_bfloat16* h2b; long long ld_h2b;
    int B, H;
    // optional fused query projection (cell 1, H == 256 == blockDim: one batch row per CTA): q[b, :] = Wq . h[b, :] + bq with
    // WqT = Wq transposed, bf16, (H, P); saves the separate 4-CTA GEMM launch of every decoder step
    const __nv_bfloat16* WqT; const float* bq; float* qout; long long ld_q; int P;
};

__global__ void __launch_bounds__(256) cell_fwd_kernel(CellFwd a) {
    pdl_wait();
    pdl_trigger();
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= a.B * a.H) return;
    const int b = idx / a.H, u = idx - b * a.H;
    const int H = a.H;
    float* g = a.G + (long long)b * 4 * H + u;
    float p0, p1, p2, p3;
    if (a.Gin) {
        p0 = p1 = p2 = p3 = 0.f;
        float q0[MAX_SPLIT], q1[MAX_SPLIT], q2[MAX_SPLIT], q3[MAX_SPLIT];
#pragma unroll
        for (int sp = 0; sp < MAX_SPLIT; ++sp) {          // every load issued before the first add
            const float* gp = a.Gin + sp * a.split_stride + (long long)b * 4 * H + u;
            const bool on = sp < a.nsplit;
            q0[sp] = on ? gp[0] : 0.f; q1[sp] = on ? gp[H] : 0.f; q2[sp] = on ? gp[2 * H] : 0.f; q3[sp] = on ? gp[3 * H] : 0.f;
        }
#pragma unroll
        for (int sp = 0; sp < MAX_SPLIT; ++sp) { p0 += q0[sp]; p1 += q1[sp]; p2 += q2[sp]; p3 += q3[sp]; }     // fixed order: deterministic
    } else {
        p0 = g[0]; p1 = g[H]; p2 = g[2 * H]; p3 = g[3 * H];
    }
    if (a.bias1) { p0 += a.bias1[u]; p1 += a.bias1[H + u]; p2 += a.bias1[2 * H + u]; p3 += a.bias1[3 * H + u]; }
    if (a.bias2) { p0 += a.bias2[u]; p1 += a.bias2[H + u]; p2 += a.bias2[2 * H + u]; p3 += a.bias2[3 * H + u]; }
    if (a.Gtab) {
        int tok;
        if (a.t == 0) tok = a.sos_idx;
        else if (a.use_gold) tok = a.y[(long long)b * a.ld_y + a.t - 1];
        else tok = a.chars_prev[b];
        if (a.tok_out && u == 0) a.tok_out[b] = tok;
        const float* tr = a.Gtab + (long long)tok * 4 * H + u;
        p0 += tr[0]; p1 += tr[H]; p2 += tr[2 * H]; p3 += tr[3 * H];
    }
    const float gi = sigmoidf_acc(p0), gf = sigmoidf_acc(p1), gg = tanhf(p2), go = sigmoidf_acc(p3);
    const float c = fmaf(gf, a.c_prev[(long long)b * a.ld_cp + u], gi * gg);
    float h = go * tanhf(c);
    if (a.mask) h *= a.mask[(long long)b * H + u];
    g[0] = gi; g[H] = gf; g[2 * H] = gg; g[3 * H] = go;
    a.c_out[(long long)b * a.ld_co + u] = c;
    a.h1[(long long)b * a.ld_h1 + u] = h;
    if (a.h2) a.h2[(long long)b * a.ld_h2 + u] = h;
    if (a.h1b) a.h1b[(long long)b * a.ld_h1b + u] = __float2bfloat16(h);
    if (a.h2b) a.h2b[(long long)b * a.ld_h2b + u] = __float2bfloat16(h);
    if (a.WqT) {
        // this CTA holds the whole h row of batch row b (H == blockDim.x == 256).  Warp w covers h units [32w, 32w+32), lane l the
        // 8 outputs [8l, 8l+8): 32 x 128-bit loads of WqT per thread, then the 8 warps' partials are added through shared memory
        __shared__ float hs[256];
        __shared__ float part[8][256];
        hs[u] = h;
        __syncthreads();
        const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (8 * l < a.P) {
#pragma unroll 8
            for (int k = 0; k < 32; ++k) {
                const int uu = w * 32 + k;
                const uint4 wv = *reinterpret_cast<const uint4*>(a.WqT + (long long)uu * a.P + 8 * l);
                const float hv = hs[uu];
                acc[0] = fmaf(__uint_as_float(wv.x << 16), hv, acc[0]); acc[1] = fmaf(__uint_as_float(wv.x & 0xffff0000u), hv, acc[1]);
                acc[2] = fmaf(__uint_as_float(wv.y << 16), hv, acc[2]); acc[3] = fmaf(__uint_as_float(wv.y & 0xffff0000u), hv, acc[3]);
                acc[4] = fmaf(__uint_as_float(wv.z << 16), hv, acc[4]); acc[5] = fmaf(__uint_as_float(wv.z & 0xffff0000u), hv, acc[5]);
                acc[6] = fmaf(__uint_as_float(wv.w << 16), hv, acc[6]); acc[7] = fmaf(__uint_as_float(wv.w & 0xffff0000u), hv, acc[7]);
            }
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) part[w][8 * l + e] = acc[e];
        __syncthreads();
        if ((int)threadIdx.x < a.P) {
            float q = a.bq[threadIdx.x];
#pragma unroll
            for (int ww = 0; ww < 8; ++ww) q += part[ww][threadIdx.x];
            a.qout[(long long)b * a.ld_q + threadIdx.x] = q;
        }
    }
}

struct CellBwd {
    float* G;                 // (B,4H) in: activated gates ; out: d(pre-activation)
    const float* dh_a; long long ld_a;   // nullable
    const float* dh_b; long long ld_b;   // nullable
    int nsplit_a, nsplit_b;              // > 1: dh_a / dh_b are un-reduced split-K partials, stride_a / stride_b floats apart
    long long stride_a, stride_b;
    const float* mask;        // (B,H) or null (dropout applied to h)
    const float* c; long long ld_c;
    const float* c_prev; long long ld_cp;
    float* dc;                // (B,H) carried in/out
    __nv_bfloat16* Gb;        // nullable (B,4H) bf16 copy of the d(pre-activation) (tensor-pipe mode)
    int first;                // 1: dc carried-in is zero
    int B, H;
};

__global__ void __launch_bounds__(256) cell_bwd_kernel(CellBwd a) {
    pdl_wait();
    pdl_trigger();
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= a.B * a.H) return;
    const int b = idx / a.H, u = idx - b * a.H;
    const int H = a.H;
    float dh = 0.f;
    {   // partial sums (at most MAX_SPLIT each): every load issued before the first add
        float pa[MAX_SPLIT], pb[MAX_SPLIT];
        const int na = a.dh_a ? (a.nsplit_a > 1 ? a.nsplit_a : 1) : 0, nb = a.dh_b ? (a.nsplit_b > 1 ? a.nsplit_b : 1) : 0;
#pragma unroll
        for (int sp = 0; sp < MAX_SPLIT; ++sp) {
            pa[sp] = sp < na ? a.dh_a[sp * a.stride_a + (long long)b * a.ld_a + u] : 0.f;
            pb[sp] = sp < nb ? a.dh_b[sp * a.stride_b + (long long)b * a.ld_b + u] : 0.f;
        }
#pragma unroll
        for (int sp = 0; sp < MAX_SPLIT; ++sp) dh += pa[sp];
#pragma unroll
        for (int sp = 0; sp < MAX_SPLIT; ++sp) dh += pb[sp];
    }
    if (a.mask) dh *= a.mask[(long long)b * H + u];
    float* g = a.G + (long long)b * 4 * H + u;
    const float gi = g[0], gf = g[H], gg = g[2 * H], go = g[3 * H];
    const float c = a.c[(long long)b * a.ld_c + u], cp = a.c_prev[(long long)b * a.ld_cp + u];
    const float dcin = a.first ? 0.f : a.dc[(long long)b * H + u];
    const float tc = tanhf(c);
    const float dct = fmaf(dh * go, 1.f - tc * tc, dcin);
    const float d0 = dct * gg * gi * (1.f - gi), d1 = dct * cp * gf * (1.f - gf);
    const float d2 = dct * gi * (1.f - gg * gg), d3 = dh * tc * go * (1.f - go);
    g[0] = d0; g[H] = d1; g[2 * H] = d2; g[3 * H] = d3;
    if (a.Gb) {
        __nv_bfloat16* gb = a.Gb + (long long)b * 4 * H + u;
        gb[0] = __float2bfloat16(d0); gb[H] = __float2bfloat16(d1); gb[2 * H] = __float2bfloat16(d2); gb[3 * H] = __float2bfloat16(d3);
    }
    a.dc[(long long)b * H + u] = dct * gf;
}

// logits[b, :] = x[b, :] . emb^T + bias ; chars[b] = argmax (first maximum, like torch.argmax)
__global__ void __launch_bounds__(256) logits_argmax_kernel(const float* __restrict__ x, long long ld_x, const float* __restrict__ emb,
                                                            const float* __restrict__ bias, float* __restrict__ logits,
                                                            long long ld_l, int* __restrict__ chars, int K, int V) {
    extern __shared__ float lsm[];   // [V]
    pdl_wait();
    pdl_trigger();
    const int b = blockIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const float* xr = x + (long long)b * ld_x;
    for (int v = w; v < V; v += 8) {
        const float* er = emb + (long long)v * K;
        float acc = 0.f;
        for (int k = lane; k < K; k += 32) acc = fmaf(xr[k], er[k], acc);
        acc = warp_sum(acc);
        if (lane == 0) lsm[v] = acc + bias[v];
    }
    __syncthreads();
    for (int v = threadIdx.x; v < V; v += 256) logits[(long long)b * ld_l + v] = lsm[v];
    if (threadIdx.x == 0) {
        int best = 0;
        float bv = lsm[0];
        for (int v = 1; v < V; ++v)
            if (lsm[v] > bv) { bv = lsm[v]; best = v; }
        chars[b] = best;
    }
}

// dGemb[v][n] = sum over (t,b) with tok[t,b] == v of dG0[t,b,n]   (deterministic gather-reduce)
__global__ void __launch_bounds__(256) token_reduce_kernel(const float* __restrict__ dG, const int* __restrict__ tok, float* __restrict__ out,
                                                           int rows, int N) {
    const int v = blockIdx.y, n = blockIdx.x * 256 + threadIdx.x;
    __shared__ int stok[256];
    float acc = 0.f;
    for (int r0 = 0; r0 < rows; r0 += 256) {
        __syncthreads();
        stok[threadIdx.x] = (r0 + threadIdx.x < rows) ? tok[r0 + threadIdx.x] : -1;
        __syncthreads();
        if (n < N) {
            const int lim = min(256, rows - r0);
            for (int i = 0; i < lim; ++i)
                if (stok[i] == v) acc += dG[(long long)(r0 + i) * N + n];
        }
    }
    if (n < N) out[(long long)v * N + n] = acc;
}

// init-force prior (reference src/models.py:326-330): block_diag of six ones((T/6+1, steps/6+1)) blocks cut to (T, steps);
// stored transposed, fm[t][t_enc], so that step t reads one contiguous row
__global__ void __launch_bounds__(256) force_mask_kernel(float* __restrict__ fm, int T, int steps) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= T * steps) return;
    const int t = i / T, te = i - t * T;
    const int a_side = T / 6 + 1, b_side = steps / 6 + 1;
    fm[i] = (te / a_side == t / b_side) ? 1.f : 0.f;
}

// two-stage deterministic column sum
constexpr int CS_ROWSPLIT = 64;
__global__ void __launch_bounds__(256) colsum_partial_kernel(const float* __restrict__ X, long long ld, int M, int N, float* __restrict__ part) {
    const int n = blockIdx.x * 32 + (threadIdx.x & 31);
    const int rl = threadIdx.x >> 5;          // 8 row lanes
    const int rs = blockIdx.y;
    __shared__ float sm[8][33];
    float acc = 0.f;
    if (n < N)
        for (int m = rs * 8 + rl; m < M; m += CS_ROWSPLIT * 8) acc += X[(long long)m * ld + n];
    sm[rl][threadIdx.x & 31] = acc;
    __syncthreads();
    if (rl == 0 && n < N) {
        float r = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) r += sm[i][threadIdx.x & 31];
        part[(long long)rs * N + n] = r;
    }
}
__global__ void __launch_bounds__(256) colsum_final_kernel(const float* __restrict__ part, int N, float* __restrict__ out, int accumulate) {
    const int n = blockIdx.x * 256 + threadIdx.x;
    if (n >= N) return;
    float r = 0.f;
    for (int i = 0; i < CS_ROWSPLIT; ++i) r += part[(long long)i * N + n];
    out[n] = accumulate ? out[n] + r : r;
}

// ------------------------------------------------------------------------------------------------------------------
// host-side helpers
// ------------------------------------------------------------------------------------------------------------------
// plain row-major GEMM helper: C[M,N](ldc) = A[M,K](lda, row-major) . op(B) + beta*C + biases
//   transB = 1: B is (N,K) row-major with row stride ldb  (C = A . B^T)
//   transB = 0: B is (K,N) row-major with row stride ldb  (C = A . B)
int gemm(cudaStream_t st, const float* A, long long lda, const float* B, long long ldb, int transB, float* C, long long ldc, int M,
         int N, int K, float beta = 0.f, const float* bias1 = nullptr, const float* bias2 = nullptr) {
    LasGemmF32 d{};
    d.A = A; d.B = B; d.C = C; d.bias1 = bias1; d.bias2 = bias2;
    d.M = M; d.N = N; d.K = K; d.batch = 1;
    d.a_m_si = lda; d.a_k_si = 1;
    if (transB) { d.b_k_si = 1; d.b_n_s = ldb; } else { d.b_k_si = ldb; d.b_n_s = 1; }
    d.c_m_si = ldc;
    d.alpha = 1.f; d.beta = beta;
    return las_gemm_f32(&d, st);
}
// C[M,N](ldc) = A^T . B where A is (K,M) row-major (lda), B is (K,N) row-major (ldb): weight-gradient form
int gemm_tn(cudaStream_t st, const float* A, long long lda, const float* B, long long ldb, float* C, long long ldc, int M, int N, int K,
            float beta = 0.f) {
    LasGemmF32 d{};
    d.A = A; d.B = B; d.C = C;
    d.M = M; d.N = N; d.K = K; d.batch = 1;
    d.a_m_si = 1; d.a_k_si = lda;
    d.b_k_si = ldb; d.b_n_s = 1;
    d.c_m_si = ldc;
    d.alpha = 1.f; d.beta = beta;
    return las_gemm_f32(&d, st);
}

// tensor-pipe forms (bf16 operands, fp32 output)
//   tc_nt: C[M,N] = A[M,K] . B[N,K]^T (+biases)     A rows stride lda, B rows stride ldb (both K contiguous)
//   tc_nn: C[M,N] = A[M,K] . B[K,N]                  B rows stride ldb (N contiguous)
//   tc_tn: C[M,N] = A[K,M]^T . B[K,N]                both reduction-major (weight gradients)
int tc_nt(cudaStream_t st, const __nv_bfloat16* A, long long lda, const __nv_bfloat16* B, long long ldb, float* C, long long ldc, int M, int N,
          int K, const float* bias1 = nullptr, const float* bias2 = nullptr) {
    LasGemmTc d{};
    d.A = A; d.B = B; d.C = C; d.bias1 = bias1; d.bias2 = bias2; d.M = M; d.N = N; d.K = K; d.a_batches = 1; d.k_batches = 1;
    d.a_s1 = lda; d.b_s1 = ldb; d.ldc = ldc;
    return las_gemm_bf16_tc(&d, st);
}
int tc_nn(cudaStream_t st, const __nv_bfloat16* A, long long lda, const __nv_bfloat16* B, long long ldb, float* C, long long ldc, int M, int N,
          int K) {
    LasGemmTc d{};
    d.A = A; d.B = B; d.C = C; d.M = M; d.N = N; d.K = K; d.a_batches = 1; d.k_batches = 1;
    d.a_s1 = lda; d.b_s1 = ldb; d.ldc = ldc; d.b_mn_major = 1;
    return las_gemm_bf16_tc(&d, st);
}
int tc_tn(cudaStream_t st, const __nv_bfloat16* A, long long lda, const __nv_bfloat16* B, long long ldb, float* C, long long ldc, int M, int N,
          int K, float* skws = nullptr, size_t skws_floats = 0) {
    LasGemmTc d{};
    d.A = A; d.B = B; d.C = C; d.M = M; d.N = N; d.K = K; d.a_batches = 1; d.k_batches = 1;
    d.a_s1 = lda; d.b_s1 = ldb; d.ldc = ldc; d.a_mn_major = 1; d.b_mn_major = 1;
    // few output tiles, long reduction (K = steps*B): split K across SMs when a workspace is available
    const int tiles = ceil_div(M, 128) * ceil_div(N, 256);
    if (skws && tiles < 74 && N % 4 == 0) {
        int sk = 148 / tiles;
        if (sk > 16) sk = 16;
        const int kiters = ceil_div(K, 64);
        if (sk > kiters) sk = kiters;
        while (sk > 1 && (size_t)sk * M * N > skws_floats) --sk;
        if (sk > 1) { d.splitk = sk; d.workspace = skws; }
    }
    return las_gemm_bf16_tc(&d, st);
}

// oh[r][v] = (tok[r] == v), v < 32: one-hot rows so that the token-wise scatter-reduce becomes a tensor-core GEMM
__global__ void __launch_bounds__(256) onehot_kernel(const int* __restrict__ tok, __nv_bfloat16* __restrict__ oh, int rows) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= rows * 32) return;
    const int r = i >> 5, v = i & 31;
    oh[i] = __float2bfloat16(tok[r] == v ? 1.f : 0.f);
}
int cast_rows(cudaStream_t st, const float* src, long long ld_src, __nv_bfloat16* dst, long long ld_dst, long long rows, int cols) {
    return las_cast_f32_to_bf16(src, ld_src, 0, 0, dst, ld_dst, rows, cols, cols, st);
}

// split factor of a per-step GEMM: as many CTAs as fit one wave, at least two 64-wide K iterations each
int pick_split(int tiles, int kiters) {
    const char* e = getenv("LAS_DEC_SPLITK");
    if (e && atoi(e) == 0) return 1;
    int sk = las_device_info()->num_sms / (tiles > 0 ? tiles : 1);
    if (sk > kiters / 2) sk = kiters / 2;
    if (sk > MAX_SPLIT) sk = MAX_SPLIT;
    return sk < 1 ? 1 : sk;
}

struct Layout {
    // float workspace offsets
    size_t Wcat0, Wcat1, Gemb, S0, S1, C0, C1, G0, G1, QC, W, W2, FM, dQC, dS0, dS1, dc0, dc1, dh1, DE, dGemb, tmpq, cs_scratch, total_f;
    // bf16 region (offsets in floats, buffers hold bf16): tensor-pipe mode only
    size_t Wcat0b, Wcat1b, Wqb, S0b, S1b, G0b, G1b, dQb, dlb, ohb, QCb, tmp32, skws, Wcat0p, Wcat1p, Gp0, Gp1, dSp0, dSp1, WqT;
    size_t skws_floats;
    // persistent decoder-step kernel (decoder_persist.cu): fp16 weights / operand rows, hand-off counters, device coin flags
    size_t Wcat0h, Wcat1h, WqTh, S0h, S1h;
    int persist;
    // int workspace offsets
    size_t tok, pctr, ugold, total_i;
    int hist, ghist;
};

bool persist_wanted(const LasSpeller* s) {
    return s->use_tc && s->kv_bf16 != 1 && las_dec_persist_fwd_supported(s->B, s->T, s->P, s->DH, s->DO, s->V, s->heads, s->init_force) != 0;
}

Layout make_layout(const LasSpeller* s) {
    Layout L{};
    const size_t B = s->B, T = s->T, P = s->P, E = s->E, DH = s->DH, DO = s->DO, V = s->V, S = s->steps, h = s->heads;
    (void)E;
    L.hist = s->training ? (int)S + 1 : 2;
    L.ghist = s->training ? (int)S : 1;
    size_t o = 0;
    auto take = [&](size_t n) { size_t r = o; o += (n + 3) & ~(size_t)3; return r; };
    L.Wcat0 = take(4 * DH * (P + DH));
    L.Wcat1 = take(4 * DO * (DH + DO));
    L.Gemb = take(V * 4 * DH);
    L.S0 = take((size_t)L.hist * B * (P + DH));
    L.S1 = take((size_t)L.hist * B * (DH + DO));
    L.C0 = take((size_t)L.hist * B * DH);
    L.C1 = take((size_t)L.hist * B * DO);
    L.G0 = take((size_t)L.ghist * B * 4 * DH);
    L.G1 = take((size_t)L.ghist * B * 4 * DO);
    L.QC = take((size_t)L.hist * B * 2 * P);
    L.W = take((size_t)L.hist * B * h * T);
    if (s->init_force) {          // second-softmax weights per step + the block-diagonal prior (steps, T)
        L.W2 = take((size_t)L.hist * B * h * T);
        L.FM = take(S * T);
    }
    if (s->training) {
        L.dQC = take((S + 1) * B * 2 * P);
        L.dS0 = take(B * (P + DH));
        L.dS1 = take(B * (DH + DO));
        L.dc0 = take(B * DH);
        L.dc1 = take(B * DO);
        L.dh1 = take(B * DO);
        L.DE = take((S + 1) * B * h * T);
        L.dGemb = take(V * 4 * DH);
        L.tmpq = take(B * DO);
        size_t maxn = 4 * DH;
        if (maxn < 2 * P) maxn = 2 * P;
        L.cs_scratch = take((size_t)CS_ROWSPLIT * maxn);
    }
    if (s->use_tc) {
        auto takeb = [&](size_t n_bf16) { return take((n_bf16 + 1) / 2 + 4); };
        L.Wcat0b = takeb(4 * DH * (P + DH));
        L.Wcat1b = takeb(4 * DO * (DH + DO));
        L.Wqb = takeb(P * DO);
        L.WqT = takeb(P * DO);
        L.Wcat0p = takeb(4 * DH * (P + DH));       // row-permuted copies for the fused LSTM epilogue (forward)
        L.Wcat1p = takeb(4 * DO * (DH + DO));
        L.S0b = takeb((size_t)L.hist * B * (P + DH));
        L.S1b = takeb((size_t)L.hist * B * (DH + DO));
        // un-reduced split-K partials of the per-step GEMMs (summed by the pointwise kernels that consume them)
        L.Gp0 = take((size_t)MAX_SPLIT * B * 4 * DH);
        L.Gp1 = take((size_t)MAX_SPLIT * B * 4 * DO);
        if (s->training) {
            L.dSp0 = take((size_t)MAX_SPLIT * B * (P + DH));
            L.dSp1 = take((size_t)MAX_SPLIT * B * (DH + DO));
        }
        if (s->training) {
            L.G0b = takeb(S * B * 4 * DH);
            L.G1b = takeb(S * B * 4 * DO);
            L.dQb = takeb((S + 1) * B * P);
            if (V <= 32) {      // padded-to-32 one-hot / dlogits operands for the token-side GEMMs
                L.dlb = takeb(S * B * 32);
                L.ohb = takeb(S * B * 32);
                L.QCb = takeb(S * B * 2 * P);
                L.tmp32 = take(32 * 4 * DH > 32 * 2 * P ? 32 * 4 * DH : 32 * 2 * P);
                L.skws_floats = (size_t)16 * 4 * DH * (P + DH);
                L.skws = take(L.skws_floats);
            }
        }
    }
    L.persist = persist_wanted(s) ? 1 : 0;
    if (L.persist) {
        auto takeh = [&](size_t n_f16) { return take((n_f16 + 1) / 2 + 4); };
        L.Wcat0h = takeh(4 * DH * (P + DH));
        L.Wcat1h = takeh(4 * DO * (DH + DO));
        L.WqTh = takeh(P * DO);
        L.S0h = takeh((size_t)L.hist * B * (P + DH));
        L.S1h = takeh((size_t)L.hist * B * (DH + DO));
    }
    L.total_f = o;
    L.tok = 0;
    size_t oi = s->training ? S * B : 4;
    oi = (oi + 31) & ~(size_t)31;                 // hand-off counters sit 128 bytes apart
    L.pctr = oi; oi += L.persist ? las_dec_persist_ctr_words((int)B) : 0;
    L.ugold = oi; oi += L.persist ? ((S + 3) & ~(size_t)3) : 0;
    L.total_i = oi;
    return L;
}

int check_speller(const LasSpeller* s) {
    LAS_CHECK_ARG(s != nullptr, "speller: null descriptor");
    LAS_CHECK_ARG(s->B >= 1 && s->T >= 1 && s->steps >= 1 && s->V >= 2, "speller: bad dims B=%d T=%d steps=%d V=%d", s->B, s->T,
                  s->steps, s->V);
    LAS_CHECK_ARG(s->E == 2 * s->P, "speller: dec_emb_dim (%d) must equal 2*att_proj_dim (%d) (tied classifier, reference "
                  "src/models.py:285-287,371)", s->E, 2 * s->P);
    LAS_CHECK_ARG(s->P % s->heads == 0, "speller: proj_dim %d %% heads %d != 0", s->P, s->heads);
    LAS_CHECK_ARG(s->P % 4 == 0 && s->DH % 4 == 0 && s->DO % 4 == 0, "speller: P/DH/DO must be multiples of 4");
    LAS_CHECK_ARG(s->emb && s->cls_b && s->w_ih0 && s->w_hh0 && s->b_ih0 && s->b_hh0 && s->w_ih1 && s->w_hh1 && s->b_ih1 && s->b_hh1 &&
                      s->wq && s->bq && s->init_query, "speller: null parameter pointer");
    LAS_CHECK_ARG(s->K && s->V_ && s->enc_lens && s->logits && s->chars && s->fws && s->iws, "speller: null input/output pointer");
    LAS_CHECK_ARG(!s->training || s->dec_y, "speller: training needs dec_y");
    LAS_CHECK_ARG(!s->use_tc || (s->P % 8 == 0 && s->DH % 8 == 0 && s->DO % 8 == 0),
                  "speller: tensor-pipe mode needs P/DH/DO multiples of 8 (TMA 16-byte strides)");
    return LAS_OK;
}

#define RC(x) do { int _rc = (x); if (_rc) return _rc; } while (0)

}  // namespace

extern "C" size_t las_colsum_scratch_floats(int N) { return (size_t)CS_ROWSPLIT * (size_t)(N > 0 ? N : 1); }

extern "C" int las_colsum_f32(const float* X, long long ld, int M, int N, float* out, int accumulate, float* scratch, void* stream) {
    LAS_CHECK_ARG(X && out && scratch && M >= 0 && N >= 1, "colsum: bad arguments");
    RC(las_set_device_of(out));
    cudaStream_t st = (cudaStream_t)stream;
    colsum_partial_kernel<<<dim3(ceil_div(N, 32), CS_ROWSPLIT), 256, 0, st>>>(X, ld, M, N, scratch);
    LAS_LAUNCH_CHECK();
    colsum_final_kernel<<<ceil_div(N, 256), 256, 0, st>>>(scratch, N, out, accumulate);
    LAS_LAUNCH_CHECK();
    return LAS_OK;
}

// standalone single-step LSTM cell pointwise (AutoRegDecoderLSTMCell.forward's API for external callers)
extern "C" int las_lstm_cell_fwd_f32(float* gates, const float* c_prev, const float* mask, float* h_out, float* c_out, int B, int H,
                                     void* stream) {
    LAS_CHECK_ARG(gates && c_prev && h_out && c_out && B >= 1 && H >= 1, "lstm_cell_fwd: bad arguments");
    RC(las_set_device_of(gates));
    CellFwd c{};
    c.G = gates; c.c_prev = c_prev; c.ld_cp = H; c.c_out = c_out; c.ld_co = H; c.mask = mask; c.h1 = h_out; c.ld_h1 = H;
    c.B = B; c.H = H;
    cell_fwd_kernel<<<ceil_div(B * H, 256), 256, 0, (cudaStream_t)stream>>>(c);
    LAS_LAUNCH_CHECK();
    return LAS_OK;
}

extern "C" int las_lstm_cell_bwd_f32(float* gates, const float* dh, const float* mask, const float* c, const float* c_prev, float* dc_io,
                                     int B, int H, void* stream) {
    LAS_CHECK_ARG(gates && dh && c && c_prev && dc_io && B >= 1 && H >= 1, "lstm_cell_bwd: bad arguments");
    RC(las_set_device_of(gates));
    CellBwd b{};
    b.G = gates; b.dh_a = dh; b.ld_a = H; b.mask = mask; b.c = c; b.ld_c = H; b.c_prev = c_prev; b.ld_cp = H; b.dc = dc_io;
    b.first = 0; b.B = B; b.H = H;
    cell_bwd_kernel<<<ceil_div(B * H, 256), 256, 0, (cudaStream_t)stream>>>(b);
    LAS_LAUNCH_CHECK();
    return LAS_OK;
}

// 1 when las_speller_fwd_f32 would run this shape as the persistent decoder-step kernel (tensor-pipe mode only); callers use it to
// pick the K / V element type (fp32 or fp16) that path accepts
extern "C" int las_speller_persistent(int B, int T, int P, int DH, int DO, int V, int heads, int init_force, int use_tc) {
    return (use_tc && las_dec_persist_fwd_supported(B, T, P, DH, DO, V, heads, init_force)) ? 1 : 0;
}

extern "C" size_t las_speller_workspace_floats(const LasSpeller* s) { return s ? make_layout(s).total_f : 0; }
extern "C" size_t las_speller_workspace_ints(const LasSpeller* s) { return s ? make_layout(s).total_i : 0; }

unsigned las_prof_mask_get();

// ---- CUDA-graph cache for the forward loop -----------------------------------------------------------------------
// The loop is ~7 launches per step (2100 at L = 300, 4200 for greedy decoding) and the host cannot enqueue them as fast
// as the GPU retires them.  The whole enqueue sequence is therefore captured once per (descriptor, coin pattern) --
// every device pointer in the descriptor is part of the key; the Python wrapper stages inputs / workspaces / outputs in a
// pointer-stable pooled buffer set (las_b200/functional.py::_SpellerSlot) -- and replayed with a single cudaGraphLaunch.  Any capture failure falls back to direct
// enqueueing.  LAS_DEC_GRAPH=0 disables it; it is also bypassed while per-kernel profiling of inner kernels is on.
namespace {
// `ident` = the bytes the key was hashed from (descriptor with the host coin pointer cleared, the coin flags, the gradient
// pointers): compared on a hash hit, so a 64-bit collision can never replay a graph recorded for other pointers / coins
struct GraphEntry { unsigned long long key; std::vector<cudaGraphExec_t> execs; unsigned long long stamp; int nlaunch; std::vector<unsigned char> ident; };

// Lets an enqueue function cut the sequence it is recording into several graphs at decoder-step boundaries.  Launching a
// 2000-node graph costs the host ~1 ms before the GPU sees its first node; when the host has no lead (the backward loop
// starts right after the host waited for the attention map of the forward loop, and greedy decoding starts cold) that is
// GPU idle time.  A short first segment lets the GPU start while the host launches the rest.  `seg == nullptr` (direct
// enqueue) and `cut_at(i)` false are no-ops.
struct GraphSeg {
    cudaStream_t cs = nullptr;
    std::vector<cudaGraphExec_t> execs;
    int cuts[4] = {0, 0, 0, 0};
    int ncuts = 0;
    bool failed = false;
    bool cut_at(int i) const {
        for (int k = 0; k < ncuts; ++k) if (cuts[k] == i) return true;
        return false;
    }
    int end_segment() {          // ends the running capture and instantiates it
        cudaGraph_t graph = nullptr;
        cudaError_t ce = cudaStreamEndCapture(cs, &graph);
        if (ce != cudaSuccess || !graph) { if (graph) cudaGraphDestroy(graph); cudaGetLastError(); failed = true; return LAS_ERR_CUDA; }
        cudaGraphExec_t exec = nullptr;
        ce = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ce != cudaSuccess || !exec) { cudaGetLastError(); failed = true; return LAS_ERR_CUDA; }
        execs.push_back(exec);
        return LAS_OK;
    }
    int cut() {
        if (end_segment() != LAS_OK) return LAS_ERR_CUDA;
        if (cudaStreamBeginCapture(cs, cudaStreamCaptureModeRelaxed) != cudaSuccess) { cudaGetLastError(); failed = true; return LAS_ERR_CUDA; }
        return LAS_OK;
    }
    void destroy() { for (auto e : execs) cudaGraphExecDestroy(e); execs.clear(); }
};
inline int seg_step(GraphSeg* seg, int i) { return (seg && seg->cut_at(i)) ? seg->cut() : LAS_OK; }
std::vector<GraphEntry> g_graphs;
unsigned long long g_graph_clock = 0;
long long g_graph_captures = 0, g_graph_replays = 0, g_graph_direct = 0;
// A capture + instantiate of a ~2000-node graph costs milliseconds; it only pays off when the same (shape, pointers, coin pattern)
// comes back.  Ragged batches or tf_rate < 1 (a new coin pattern per batch) never repeat: after a short streak of misses, keys that
// were not seen recently are enqueued directly instead of being captured.
unsigned long long g_recent_keys[16];
int g_recent_n = 0, g_miss_streak = 0;
cudaStream_t g_capture_stream[64];
bool g_capture_stream_ok[64];

unsigned long long fnv1a(const void* p, size_t n, unsigned long long h) {
    const unsigned char* c = (const unsigned char*)p;
    for (size_t i = 0; i < n; ++i) { h ^= c[i]; h *= 1099511628211ULL; }
    return h;
}
int speller_fwd_enqueue(const LasSpeller* s, const Layout& L, cudaStream_t st, GraphSeg* seg);
int speller_fwd_persist(const LasSpeller* s, const Layout& L, cudaStream_t st);
}  // namespace

namespace {
int speller_bwd_enqueue(const LasSpeller* s, const LasSpellerGrads* g, const Layout& L, cudaStream_t st, GraphSeg* seg, int phases);
std::mutex g_graph_mu;

// Runs `enqueue(stream)` through the graph cache: replay when `key` is known, else capture on a side stream, instantiate,
// remember (LRU of 8) and launch.  Any capture failure (and LAS_DEC_GRAPH=0, or per-kernel profiling of the inner kernels)
// falls back to enqueueing directly on `st`.
template <class F>
int run_graph_cached(unsigned long long key, const std::vector<unsigned char>& ident, cudaStream_t st, F enqueue) {
    const char* genv = getenv("LAS_DEC_GRAPH");
    const bool inner_prof = (las_prof_mask_get() & ((1u << LAS_PROF_GEMM_OTHER) | (1u << LAS_PROF_ATTN_FWD) | (1u << LAS_PROF_ATTN_BWD))) != 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if ((genv && atoi(genv) == 0) || inner_prof || dev < 0 || dev >= 64) return enqueue(st, (GraphSeg*)nullptr);
    key = fnv1a(&dev, sizeof(dev), key);
    std::lock_guard<std::mutex> lk(g_graph_mu);
    for (auto& e : g_graphs)
        if (e.key == key && e.ident == ident) {
            e.stamp = ++g_graph_clock;
            ++g_graph_replays;
            g_miss_streak = 0;
            for (auto x : e.execs) LAS_CUDA(cudaGraphLaunch(x, st));
            las_count_launch(e.nlaunch);          // the replay runs the same kernels the capture recorded
            return LAS_OK;
        }
    {
        bool seen = false;
        for (int i = 0; i < g_recent_n; ++i) seen = seen || g_recent_keys[i] == key;
        if (g_recent_n < 16) g_recent_keys[g_recent_n++] = key;
        else { for (int i = 1; i < 16; ++i) g_recent_keys[i - 1] = g_recent_keys[i]; g_recent_keys[15] = key; }
        ++g_miss_streak;
        if (g_miss_streak > 3 && !seen) { ++g_graph_direct; return enqueue(st, (GraphSeg*)nullptr); }
    }
    if (!g_capture_stream_ok[dev]) {
        if (cudaStreamCreateWithFlags(&g_capture_stream[dev], cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); return enqueue(st, (GraphSeg*)nullptr); }
        g_capture_stream_ok[dev] = true;
    }
    GraphSeg seg;
    seg.cs = g_capture_stream[dev];
    {   // decoder steps (counted from the first one enqueued) before which the graph is cut; LAS_DEC_GRAPH_CUTS="a,b,..", "0" = one graph
        const char* cenv = getenv("LAS_DEC_GRAPH_CUTS");
        const char* c = cenv ? cenv : "16,64";
        while (*c && seg.ncuts < 4) {
            const int v = atoi(c);
            if (v > 0) seg.cuts[seg.ncuts++] = v;
            while (*c && *c != ',') ++c;
            if (*c == ',') ++c;
        }
    }
    if (cudaStreamBeginCapture(seg.cs, cudaStreamCaptureModeRelaxed) != cudaSuccess) { cudaGetLastError(); return enqueue(st, (GraphSeg*)nullptr); }
    const long long launches_before = las_launch_count();
    int rc = enqueue(seg.cs, &seg);
    if (seg.failed) {            // a cut failed: the stream may or may not still be capturing
        cudaStreamCaptureStatus cst = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(seg.cs, &cst) == cudaSuccess && cst != cudaStreamCaptureStatusNone) { cudaGraph_t g2 = nullptr; cudaStreamEndCapture(seg.cs, &g2); if (g2) cudaGraphDestroy(g2); }
        cudaGetLastError();
        seg.destroy();
        return enqueue(st, (GraphSeg*)nullptr);
    }
    if (rc != LAS_OK) {
        cudaGraph_t g2 = nullptr;
        cudaStreamEndCapture(seg.cs, &g2);
        if (g2) cudaGraphDestroy(g2);
        cudaGetLastError();
        seg.destroy();
        return rc;
    }
    if (seg.end_segment() != LAS_OK) { seg.destroy(); return enqueue(st, (GraphSeg*)nullptr); }
    if (g_graphs.size() >= 8) {           // evict the least recently used
        size_t lru = 0;
        for (size_t i = 1; i < g_graphs.size(); ++i)
            if (g_graphs[i].stamp < g_graphs[lru].stamp) lru = i;
        for (auto x : g_graphs[lru].execs) cudaGraphExecDestroy(x);
        g_graphs.erase(g_graphs.begin() + lru);
    }
    g_graphs.push_back({key, seg.execs, ++g_graph_clock, (int)(las_launch_count() - launches_before), ident});
    ++g_graph_captures;
    for (auto x : seg.execs) LAS_CUDA(cudaGraphLaunch(x, st));
    return LAS_OK;
}
}  // namespace

extern "C" void las_speller_graph_stats(long long* captures, long long* replays) {
    if (captures) *captures = g_graph_captures;
    if (replays) *replays = g_graph_replays;
}

extern "C" int las_speller_fwd_f32(const LasSpeller* s, void* stream) {
    RC(check_speller(s));
    const Layout L = make_layout(s);
    if (s->fws_floats < L.total_f || s->iws_ints < L.total_i) {
        las_set_error("speller_fwd: workspace too small (%zu/%zu floats, %zu/%zu ints)", s->fws_floats, L.total_f, s->iws_ints, L.total_i);
        return LAS_ERR_WORKSPACE;
    }
    RC(las_set_device_of(s->fws));
    cudaStream_t st = (cudaStream_t)stream;
    LasProfScope prof(LAS_PROF_SPELLER_FWD, stream, (double)s->steps);
    // key: every field of the descriptor except the HOST coin pointer (a fresh host array per call), whose contents are hashed
    LasSpeller kd;
    memcpy(&kd, s, sizeof(LasSpeller));
    kd.use_gold_host = nullptr;
    std::vector<unsigned char> ident((const unsigned char*)&kd, (const unsigned char*)&kd + sizeof(LasSpeller));
    if (s->use_gold_host) ident.insert(ident.end(), s->use_gold_host, s->use_gold_host + s->steps);
    const unsigned long long key = fnv1a(ident.data(), ident.size(), 1469598103934665603ULL);
    // persistent decoder-step kernel: one cooperative launch for the whole loop (plus a handful of weight-preparation kernels),
    // enqueued directly -- nothing to amortise with a graph, and no graph key that depends on the coin pattern
    if (L.persist) {
        const int rc = speller_fwd_persist(s, L, st);
        if (rc == LAS_OK || s->kv_bf16 == 2) return rc;          // fp16 K / V only exist for the persistent kernel
        cudaGetLastError();      // e.g. the cooperative launch was refused (device shared through MPS): the launch-per-stage loop still runs
    }
    return run_graph_cached(key, ident, st, [&](cudaStream_t q, GraphSeg* seg) { return speller_fwd_enqueue(s, L, q, seg); });
}

namespace {
int speller_fwd_persist(const LasSpeller* s, const Layout& L, cudaStream_t st) {
    const int B = s->B, T = s->T, P = s->P, E = s->E, DH = s->DH, DO = s->DO, V = s->V, S = s->steps;
    const int K0 = P + DH, K1 = DH + DO;
    float* f = s->fws;
    float *Wcat0 = f + L.Wcat0, *Wcat1 = f + L.Wcat1, *Gemb = f + L.Gemb, *C0 = f + L.C0, *C1 = f + L.C1, *QC = f + L.QC;
    const size_t fsz = sizeof(float);
    __nv_bfloat16 *Wcat0b = (__nv_bfloat16*)(f + L.Wcat0b), *Wcat1b = (__nv_bfloat16*)(f + L.Wcat1b), *Wqb = (__nv_bfloat16*)(f + L.Wqb),
                  *S0b = (__nv_bfloat16*)(f + L.S0b), *S1b = (__nv_bfloat16*)(f + L.S1b);
    __half *Wcat0h = (__half*)(f + L.Wcat0h), *Wcat1h = (__half*)(f + L.Wcat1h), *WqTh = (__half*)(f + L.WqTh), *S0h = (__half*)(f + L.S0h),
           *S1h = (__half*)(f + L.S1h);
    // packed weights: [W_ih0[:, E:] | W_hh0], [W_ih1 | W_hh1]
    LAS_CUDA(cudaMemcpy2DAsync(Wcat0, K0 * fsz, s->w_ih0 + E, (size_t)(E + P) * fsz, P * fsz, 4 * DH, cudaMemcpyDeviceToDevice, st));
    LAS_CUDA(cudaMemcpy2DAsync(Wcat0 + P, K0 * fsz, s->w_hh0, DH * fsz, DH * fsz, 4 * DH, cudaMemcpyDeviceToDevice, st));
    LAS_CUDA(cudaMemcpy2DAsync(Wcat1, K1 * fsz, s->w_ih1, DH * fsz, DH * fsz, 4 * DO, cudaMemcpyDeviceToDevice, st));
    LAS_CUDA(cudaMemcpy2DAsync(Wcat1 + DH, K1 * fsz, s->w_hh1, DO * fsz, DO * fsz, 4 * DO, cudaMemcpyDeviceToDevice, st));
    RC(las_cast_f32_to_f16(Wcat0, K0, 0, 0, Wcat0h, K0, 4 * DH, K0, K0, st));
    RC(las_cast_f32_to_f16(Wcat1, K1, 0, 0, Wcat1h, K1, 4 * DO, K1, K1, st));
    RC(las_transpose_cast_f16(s->wq, WqTh, P, DO, st));                          // (P, DO) fp32 -> (DO, P) fp16
    if (s->training) {                                                           // backward's bf16 operands
        RC(cast_rows(st, Wcat0, K0, Wcat0b, K0, 4 * DH, K0));
        RC(cast_rows(st, Wcat1, K1, Wcat1b, K1, 4 * DO, K1));
        RC(cast_rows(st, s->wq, DO, Wqb, DO, P, DO));
        LAS_CUDA(cudaMemset2DAsync(S0b + P, K0 * 2, 0, DH * 2, B, st));          // h0_{-1} = 0
        LAS_CUDA(cudaMemset2DAsync(S1b + DH, K1 * 2, 0, DO * 2, B, st));         // h1_{-1} = 0
    }
    LAS_CUDA(cudaMemset2DAsync(S0h + P, K0 * 2, 0, DH * 2, B, st));
    LAS_CUDA(cudaMemset2DAsync(S1h + DH, K1 * 2, 0, DO * 2, B, st));
    // embedding-side gate table (+ both cell-0 biases); zero initial cell states (src/models.py:275-281)
    RC(gemm(st, s->emb, E, s->w_ih0, E + P, 1, Gemb, 4 * DH, V, 4 * DH, E, 0.f, s->b_ih0, s->b_hh0));
    LAS_CUDA(cudaMemsetAsync(C0, 0, (size_t)B * DH * fsz, st));
    LAS_CUDA(cudaMemsetAsync(C1, 0, (size_t)B * DO * fsz, st));
    unsigned* ctr = (unsigned*)(s->iws + L.pctr);
    const size_t ctr_words = las_dec_persist_ctr_words(B);
    LAS_CUDA(cudaMemsetAsync(ctr, 0, ctr_words * sizeof(unsigned), st));
    bool all_gold = s->training != 0;
    if (s->training)
        for (int t = 1; t < S; ++t) all_gold = all_gold && s->use_gold_host && s->use_gold_host[t];
    int* ugold = s->iws + L.ugold;
    if (s->training && !all_gold) {
        std::vector<int> ug((size_t)S, 0);
        for (int t = 1; t < S; ++t) ug[t] = (s->use_gold_host && s->use_gold_host[t]) ? 1 : 0;
        LAS_CUDA(cudaMemcpyAsync(ugold, ug.data(), (size_t)S * sizeof(int), cudaMemcpyHostToDevice, st));   // pageable source: staged before the call returns
    }
    LasDecPersistFwd a{};
    a.B = B; a.T = T; a.P = P; a.DH = DH; a.DO = DO; a.V = V; a.steps = S; a.training = s->training; a.sos_idx = s->sos_idx;
    a.hist = L.hist; a.ghist = L.ghist;
    a.per_step_logits = all_gold ? 0 : 1;
    a.kv16 = s->kv_bf16 == 2 ? 1 : 0;
    a.scale = sqrtf((float)(P / s->heads));
    a.emb = s->emb; a.cls_b = s->cls_b; a.bq = s->bq; a.b_ih1 = s->b_ih1; a.b_hh1 = s->b_hh1; a.init_query = s->init_query;
    a.Gemb = Gemb; a.W0 = Wcat0h; a.W1 = Wcat1h; a.WqT = WqTh;
    a.K = s->K; a.Vv = s->V_; a.enc_lens = s->enc_lens;
    a.y = s->training ? s->dec_y : nullptr; a.ld_y = s->ld_y;
    a.use_gold = (s->training && !all_gold) ? ugold : nullptr;
    a.drop0 = s->drop0; a.drop1 = s->drop1;
    a.S0h = S0h; a.S1h = S1h;
    a.S0b = s->training ? S0b : nullptr; a.S1b = s->training ? S1b : nullptr;
    a.C0 = C0; a.C1 = C1; a.G0 = f + L.G0; a.G1 = f + L.G1; a.QC = QC; a.W = f + L.W;
    a.att0 = s->att0; a.logits = s->logits; a.chars = s->chars;
    a.tok = s->training ? s->iws + L.tok : nullptr;
    a.ctr = ctr; a.err = ctr + (ctr_words - 32);
    RC(las_dec_persist_fwd_launch(&a, st));
    if (all_gold) {
        // all classifier rows at once: row m = t*B + b of QC[1..S] -> logits[b, t, :]
        LasGemmF32 d{};
        d.A = QC + (size_t)B * 2 * P; d.B = s->emb; d.C = s->logits; d.bias1 = s->cls_b;
        d.M = S * B; d.N = V; d.K = 2 * P; d.batch = 1;
        d.a_m_si = 2 * P; d.a_k_si = 1; d.b_k_si = 1; d.b_n_s = E;
        d.c_m_inner = B; d.c_m_so = V; d.c_m_si = (long long)S * V;
        d.alpha = 1.f; d.beta = 0.f;
        RC(las_gemm_f32(&d, st));
    }
    return LAS_OK;
}

int speller_fwd_enqueue(const LasSpeller* s, const Layout& L, cudaStream_t st, GraphSeg* seg) {
    void* stream = (void*)st;
    (void)stream;
    const int B = s->B, T = s->T, P = s->P, E = s->E, DH = s->DH, DO = s->DO, V = s->V, S = s->steps, heads = s->heads;
    const int K0 = P + DH, K1 = DH + DO;
    float* f = s->fws;
    float *Wcat0 = f + L.Wcat0, *Wcat1 = f + L.Wcat1, *Gemb = f + L.Gemb, *S0 = f + L.S0, *S1 = f + L.S1, *C0 = f + L.C0, *C1 = f + L.C1,
          *G0 = f + L.G0, *G1 = f + L.G1, *QC = f + L.QC, *W = f + L.W;
    int* tok = s->iws + L.tok;
    const size_t fsz = sizeof(float);
    const bool tc = s->use_tc != 0;
    __nv_bfloat16 *Wcat0b = (__nv_bfloat16*)(f + L.Wcat0b), *Wcat1b = (__nv_bfloat16*)(f + L.Wcat1b), *Wqb = (__nv_bfloat16*)(f + L.Wqb),
                  *S0b = (__nv_bfloat16*)(f + L.S0b), *S1b = (__nv_bfloat16*)(f + L.S1b);

    // packed weights
    LAS_CUDA(cudaMemcpy2DAsync(Wcat0, K0 * fsz, s->w_ih0 + E, (size_t)(E + P) * fsz, P * fsz, 4 * DH, cudaMemcpyDeviceToDevice, st));
    LAS_CUDA(cudaMemcpy2DAsync(Wcat0 + P, K0 * fsz, s->w_hh0, DH * fsz, DH * fsz, 4 * DH, cudaMemcpyDeviceToDevice, st));
    LAS_CUDA(cudaMemcpy2DAsync(Wcat1, K1 * fsz, s->w_ih1, DH * fsz, DH * fsz, 4 * DO, cudaMemcpyDeviceToDevice, st));
    LAS_CUDA(cudaMemcpy2DAsync(Wcat1 + DH, K1 * fsz, s->w_hh1, DO * fsz, DO * fsz, 4 * DO, cudaMemcpyDeviceToDevice, st));
    if (tc) {
        RC(cast_rows(st, Wcat0, K0, Wcat0b, K0, 4 * DH, K0));
        RC(cast_rows(st, Wcat1, K1, Wcat1b, K1, 4 * DO, K1));
        RC(cast_rows(st, s->wq, DO, Wqb, DO, P, DO));
        RC(las_transpose_cast_bf16(s->wq, f + L.WqT, 1, P, DO, st));          // WqT (DO, P) bf16 for the query projection fused into cell 1
        LAS_CUDA(cudaMemset2DAsync(S0b + P, K0 * 2, 0, DH * 2, B, st));       // h0_{-1} = 0 (bf16)
        LAS_CUDA(cudaMemset2DAsync(S1b + DH, K1 * 2, 0, DO * 2, B, st));      // h1_{-1} = 0
    }
    PlanBuf pl0, pl1, plq;
    // LSTM pointwise fused into the GEMM epilogue: measured SLOWER than GEMM + coalesced cell kernel (thread-per-row epilogue
    // accesses are uncoalesced), so it is opt-in (LAS_DEC_FUSE=1) until the epilogue is staged through shared memory
    const char* fuse_env = getenv("LAS_DEC_FUSE");
    const bool fuse = tc && DH % 16 == 0 && DO % 16 == 0 && fuse_env && atoi(fuse_env) == 1;
    if (tc) {
        LAS_CHECK_ARG(las_tc_plan_bytes() <= sizeof(PlanBuf), "speller: plan buffer too small");
        __nv_bfloat16 *Wcat0p = (__nv_bfloat16*)(f + L.Wcat0p), *Wcat1p = (__nv_bfloat16*)(f + L.Wcat1p);
        if (fuse) {
            RC(las_permute_cast_lstm_rows(Wcat0, K0, Wcat0p, DH, K0, st));
            RC(las_permute_cast_lstm_rows(Wcat1, K1, Wcat1p, DO, K1, st));
        }
        RC(las_tc_plan_make(&pl0, S0b, fuse ? Wcat0p : Wcat0b, B, 4 * DH, K0, L.hist, K0, (long long)B * K0, K0, 0));
        RC(las_tc_plan_make(&pl1, S1b, fuse ? Wcat1p : Wcat1b, B, 4 * DO, K1, L.hist, K1, (long long)B * K1, K1, 0));
        RC(las_tc_plan_make(&plq, S1b + DH, Wqb, B, P, DO, L.hist, K1, (long long)B * K1, DO, 0));
    }
    // split-K of the two cell GEMMs (un-reduced partials, summed by cell_fwd_kernel): 32 / 16 output tiles cannot fill 148 SMs
    const int sk0 = (tc && !fuse) ? pick_split(las_tc_plan_tiles(&pl0), las_tc_plan_kiters(&pl0)) : 1;
    const int sk1 = (tc && !fuse) ? pick_split(las_tc_plan_tiles(&pl1), las_tc_plan_kiters(&pl1)) : 1;
    float *Gp0 = f + L.Gp0, *Gp1 = f + L.Gp1;
    // query projection fused into the cell-1 kernel (one CTA = one batch row needs DO == 256 == its block size)
    const char* fq_env = getenv("LAS_DEC_FUSEQ");
    const bool fuse_q = tc && !fuse && DO == 256 && P <= 256 && P % 8 == 0 && !(fq_env && atoi(fq_env) == 0);
    // embedding-side gate table (+ both cell-0 biases)
    RC(gemm(st, s->emb, E, s->w_ih0, E + P, 1, Gemb, 4 * DH, V, 4 * DH, E, 0.f, s->b_ih0, s->b_hh0));
    // zero initial states (init_hiddens are always zero: reference src/models.py:275-281)
    LAS_CUDA(cudaMemset2DAsync(S0 + P, K0 * fsz, 0, DH * fsz, B, st));      // h0_{-1}
    LAS_CUDA(cudaMemset2DAsync(S1 + DH, K1 * fsz, 0, DO * fsz, B, st));     // h1_{-1}
    LAS_CUDA(cudaMemsetAsync(C0, 0, (size_t)B * DH * fsz, st));
    LAS_CUDA(cudaMemsetAsync(C1, 0, (size_t)B * DO * fsz, st));
    // initial query (src/models.py:345-346): q_init = query_map(init_query) for every row
    {
        LasGemmF32 d{};
        d.A = s->init_query; d.B = s->wq; d.C = QC; d.bias1 = s->bq;
        d.M = B; d.N = P; d.K = DO; d.batch = 1;
        d.a_m_si = 0; d.a_k_si = 1; d.b_k_si = 1; d.b_n_s = DO; d.c_m_si = 2 * P;
        d.alpha = 1.f; d.beta = 0.f;
        RC(las_gemm_f32(&d, st));
    }
    LasAttnStep at{};
    at.K = s->K; at.V = s->V_; at.lens = s->enc_lens; at.B = B; at.T = T; at.P = P; at.heads = heads;
    at.scale = sqrtf((float)(P / heads));
    at.kv_bf16 = s->kv_bf16;
    at.ld_q = 2 * P; at.ld_ctx = 2 * P; at.ld_ctx2 = K0; at.ld_w = T;
    at.q = QC; at.ctx = QC + P; at.ctx2 = S0; at.w = W; at.w_b0 = s->att0;
    at.ctx2_bf16 = tc ? (void*)S0b : nullptr; at.ld_ctx2_bf16 = K0;
    RC(las_attn_step_fwd_f32(&at, st));
    float *W2 = f + L.W2, *FM = f + L.FM;
    if (s->init_force) {
        force_mask_kernel<<<ceil_div(S * T, 256), 256, 0, st>>>(FM, T, S);
        LAS_LAUNCH_CHECK();
        // the initial attention has no prior: its weights double as "second-softmax" weights for the batched dV GEMM
        if (s->training) LAS_CUDA(cudaMemcpyAsync(W2, W, (size_t)B * heads * T * fsz, cudaMemcpyDeviceToDevice, st));
    }

    bool all_gold = s->training != 0;
    if (s->training)
        for (int t = 1; t < S; ++t) all_gold = all_gold && s->use_gold_host && s->use_gold_host[t];
    const bool per_step_logits = !all_gold;

    LasPdlScope pdl_scope;       // the per-step kernels below overlap their launch / prologue with the predecessor's tail
    for (int t = 0; t < S; ++t) {
        RC(seg_step(seg, t));
        const int r = t % L.hist, rn = (t + 1) % L.hist, rg = t % L.ghist;
        float* S0r = S0 + (size_t)r * B * K0;  float* S0n = S0 + (size_t)rn * B * K0;
        float* S1r = S1 + (size_t)r * B * K1;  float* S1n = S1 + (size_t)rn * B * K1;
        float* G0r = G0 + (size_t)rg * B * 4 * DH;  float* G1r = G1 + (size_t)rg * B * 4 * DO;
        float* QCn = QC + (size_t)rn * B * 2 * P;
        __nv_bfloat16* S0rb = S0b + (size_t)r * B * K0;  __nv_bfloat16* S0nb = S0b + (size_t)rn * B * K0;
        __nv_bfloat16* S1rb = S1b + (size_t)r * B * K1;  __nv_bfloat16* S1nb = S1b + (size_t)rn * B * K1;
        // cell 0 and cell 1
        if (fuse) {
            LasLstmEpi e0{};
            e0.H = DH; e0.tab = Gemb; e0.y = s->dec_y; e0.ld_y = s->ld_y;
            e0.chars_prev = (t > 0) ? s->chars + (size_t)(t - 1) * B : nullptr;
            e0.tok_out = s->training ? tok + (size_t)t * B : nullptr;
            e0.t = t; e0.sos_idx = s->sos_idx;
            e0.use_gold = (s->training && t > 0 && s->use_gold_host && s->use_gold_host[t]) ? 1 : 0;
            e0.c_prev = C0 + (size_t)r * B * DH; e0.ld_cp = DH; e0.c_out = C0 + (size_t)rn * B * DH; e0.ld_co = DH;
            e0.mask = s->drop0 ? s->drop0 + (size_t)t * B * DH : nullptr;
            e0.G = G0r;
            e0.h1 = S0n + P; e0.ld_h1 = K0; e0.h2 = S1r; e0.ld_h2 = K1;
            e0.h1b = S0nb + P; e0.ld_h1b = K0; e0.h2b = S1rb; e0.ld_h2b = K1;
            RC(las_tc_plan_launch_lstm(&pl0, r, &e0, st));
            LasLstmEpi e1{};
            e1.H = DO; e1.bias1 = s->b_ih1; e1.bias2 = s->b_hh1; e1.t = t;
            e1.c_prev = C1 + (size_t)r * B * DO; e1.ld_cp = DO; e1.c_out = C1 + (size_t)rn * B * DO; e1.ld_co = DO;
            e1.mask = s->drop1 ? s->drop1 + (size_t)t * B * DO : nullptr;
            e1.G = G1r;
            e1.h1 = S1n + DH; e1.ld_h1 = K1; e1.h1b = S1nb + DH; e1.ld_h1b = K1;
            RC(las_tc_plan_launch_lstm(&pl1, r, &e1, st));
        } else {
        // cell 0
            if (tc) RC(las_tc_plan_launch_split(&pl0, r, Gp0, 4 * DH, sk0, st));
            else RC(gemm(st, S0r, K0, Wcat0, K0, 1, G0r, 4 * DH, B, 4 * DH, K0));
            CellFwd c0{};
            if (tc) { c0.Gin = Gp0; c0.nsplit = sk0; c0.split_stride = (long long)B * 4 * DH; }
            c0.G = G0r; c0.Gtab = Gemb; c0.y = s->dec_y; c0.ld_y = s->ld_y;
            c0.chars_prev = (t > 0) ? s->chars + (size_t)(t - 1) * B : nullptr;
            c0.tok_out = s->training ? tok + (size_t)t * B : nullptr;
            c0.t = t; c0.sos_idx = s->sos_idx;
            c0.use_gold = (s->training && t > 0 && s->use_gold_host && s->use_gold_host[t]) ? 1 : 0;
            c0.c_prev = C0 + (size_t)r * B * DH; c0.ld_cp = DH;
            c0.c_out = C0 + (size_t)rn * B * DH; c0.ld_co = DH;
            c0.mask = s->drop0 ? s->drop0 + (size_t)t * B * DH : nullptr;
            c0.h1 = S0n + P; c0.ld_h1 = K0;      // recurrent slot of the next step's cell-0 row
            c0.h2 = S1r; c0.ld_h2 = K1;          // input slot of this step's cell-1 row
            if (tc) { c0.h1b = S0nb + P; c0.ld_h1b = K0; c0.h2b = S1rb; c0.ld_h2b = K1; }
            c0.B = B; c0.H = DH;
            LAS_CUDA(las_launch(cell_fwd_kernel, dim3(ceil_div(B * DH, 256)), dim3(256), 0, st, c0));
            LAS_LAUNCH_CHECK();
            // cell 1
            if (tc) RC(las_tc_plan_launch_split(&pl1, r, Gp1, 4 * DO, sk1, st));
            else RC(gemm(st, S1r, K1, Wcat1, K1, 1, G1r, 4 * DO, B, 4 * DO, K1, 0.f, s->b_ih1, s->b_hh1));
            CellFwd c1{};
            if (tc) { c1.Gin = Gp1; c1.nsplit = sk1; c1.split_stride = (long long)B * 4 * DO; c1.bias1 = s->b_ih1; c1.bias2 = s->b_hh1; }
            c1.G = G1r; c1.Gtab = nullptr; c1.t = t;
            c1.c_prev = C1 + (size_t)r * B * DO; c1.ld_cp = DO;
            c1.c_out = C1 + (size_t)rn * B * DO; c1.ld_co = DO;
            c1.mask = s->drop1 ? s->drop1 + (size_t)t * B * DO : nullptr;
            c1.h1 = S1n + DH; c1.ld_h1 = K1; c1.h2 = nullptr;
            if (tc) { c1.h1b = S1nb + DH; c1.ld_h1b = K1; }
            c1.B = B; c1.H = DO;
            if (fuse_q) { c1.WqT = (const __nv_bfloat16*)(f + L.WqT); c1.bq = s->bq; c1.qout = QCn; c1.ld_q = 2 * P; c1.P = P; }
            LAS_CUDA(las_launch(cell_fwd_kernel, dim3(ceil_div(B * DO, 256)), dim3(256), 0, st, c1));
            LAS_LAUNCH_CHECK();
        }
        // query projection into QC[t+1][:, :P]
        if (fuse_q) { /* q was written by cell 1 */ }
        else if (tc) RC(las_tc_plan_launch(&plq, rn, QCn, 2 * P, s->bq, nullptr, st));
        else RC(gemm(st, S1n + DH, K1, s->wq, DO, 1, QCn, 2 * P, B, P, DO, 0.f, s->bq));
        // attention: context into QC[t+1][:, P:] and into the next cell-0 row
        at.q = QCn; at.ctx = QCn + P; at.ctx2 = S0n; at.w = W + (size_t)rn * B * heads * T;
        at.ctx2_bf16 = tc ? (void*)S0nb : nullptr;
        at.w_b0 = s->att0 ? s->att0 + (size_t)(t + 1) * heads * T : nullptr;
        if (s->init_force) { at.fmask = FM + (size_t)t * T; at.ld_fmask = 0; at.w2 = W2 + (size_t)rn * B * heads * T; }
        RC(las_attn_step_fwd_f32(&at, st));
        if (per_step_logits) {
            LAS_CUDA(las_launch(logits_argmax_kernel, dim3(B), dim3(256), V * sizeof(float), st, (const float*)QCn, (long long)(2 * P), s->emb,
                                s->cls_b, s->logits + (size_t)t * V, (long long)S * V, s->chars + (size_t)t * B, 2 * P, V));
            LAS_LAUNCH_CHECK();
        }
    }
    if (!per_step_logits) {
        // all classifier rows at once: row m = t*B + b of QC[1..S] -> logits[b, t, :]
        LasGemmF32 d{};
        d.A = QC + (size_t)B * 2 * P; d.B = s->emb; d.C = s->logits; d.bias1 = s->cls_b;
        d.M = S * B; d.N = V; d.K = 2 * P; d.batch = 1;
        d.a_m_si = 2 * P; d.a_k_si = 1; d.b_k_si = 1; d.b_n_s = E;
        d.c_m_inner = B; d.c_m_so = V; d.c_m_si = (long long)S * V;
        d.alpha = 1.f; d.beta = 0.f;
        RC(las_gemm_f32(&d, st));
    }
    return LAS_OK;
}

}  // namespace

// phases: bit 0 = the backward time loop + the initial attention step + dK / dV (what the encoder's backward waits for),
//         bit 1 = the batched parameter gradients (they feed nothing but the optimizer: a caller may issue them on another stream
//                 once phase 1 has been enqueued -- las_b200.functional.SpellerFunction runs them beside the top encoder layer's BPTT
//                 kernel).  Each phase set is its own cached CUDA graph.
extern "C" int las_speller_bwd_phases_f32(const LasSpeller* s, const LasSpellerGrads* g, int phases, void* stream) {
    RC(check_speller(s));
    LAS_CHECK_ARG(s->training, "speller_bwd: forward was not run in training mode");
    LAS_CHECK_ARG(phases >= 1 && phases <= 3, "speller_bwd: phases must be 1 (loop), 2 (parameter gradients) or 3 (both)");
    LAS_CHECK_ARG(g && g->dlogits && g->d_emb && g->d_cls_b && g->d_w_ih0 && g->d_w_hh0 && g->d_b_ih0 && g->d_b_hh0 && g->d_w_ih1 &&
                      g->d_w_hh1 && g->d_b_ih1 && g->d_b_hh1 && g->d_wq && g->d_bq && g->d_init_query && g->dK && g->dV,
                  "speller_bwd: null gradient pointer");
    const Layout L = make_layout(s);
    if (s->fws_floats < L.total_f || s->iws_ints < L.total_i) {
        las_set_error("speller_bwd: workspace too small");
        return LAS_ERR_WORKSPACE;
    }
    RC(las_set_device_of(s->fws));
    cudaStream_t st = (cudaStream_t)stream;
    // the bench's decoder-backward time is the loop; the parameter gradients issued alone (beside a BPTT kernel) are not counted as one
    std::unique_ptr<LasProfScope> prof;
    if (phases & 1) prof.reset(new LasProfScope(LAS_PROF_SPELLER_BWD, stream, (double)s->steps));
    // same graph cache as the forward loop: key = descriptor (coin contents hashed) + every gradient pointer + the phases
    LasSpeller kd;
    memcpy(&kd, s, sizeof(LasSpeller));
    kd.use_gold_host = nullptr;
    std::vector<unsigned char> ident((const unsigned char*)&kd, (const unsigned char*)&kd + sizeof(LasSpeller));
    if (s->use_gold_host) ident.insert(ident.end(), s->use_gold_host, s->use_gold_host + s->steps);
    ident.insert(ident.end(), (const unsigned char*)g, (const unsigned char*)g + sizeof(LasSpellerGrads));
    for (const char* name : {"LAS_BWD_FUSE_TAIL", "LAS_ATTN_SPLIT", "LAS_BWD_ATT_TC"}) {       // tuning switches that change which kernels the graph holds
        const char* e = getenv(name);
        ident.push_back((unsigned char)((e && *e) ? *e : 0));
    }
    ident.push_back((unsigned char)phases);
    const unsigned long long key = fnv1a(ident.data(), ident.size(), 7809847782465536322ULL);
    return run_graph_cached(key, ident, st, [&](cudaStream_t q, GraphSeg* seg) { return speller_bwd_enqueue(s, g, L, q, seg, phases); });
}

extern "C" int las_speller_bwd_f32(const LasSpeller* s, const LasSpellerGrads* g, void* stream) {
    return las_speller_bwd_phases_f32(s, g, 3, stream);
}

namespace {
int speller_bwd_enqueue(const LasSpeller* s, const LasSpellerGrads* g, const Layout& L, cudaStream_t st, GraphSeg* seg, int phases) {
    const int B = s->B, T = s->T, P = s->P, E = s->E, DH = s->DH, DO = s->DO, V = s->V, S = s->steps, heads = s->heads;
    const int K0 = P + DH, K1 = DH + DO, d_head = P / heads;
    float* f = s->fws;
    float *Wcat0 = f + L.Wcat0, *Wcat1 = f + L.Wcat1, *S0 = f + L.S0, *S1 = f + L.S1, *C0 = f + L.C0, *C1 = f + L.C1, *G0 = f + L.G0,
          *G1 = f + L.G1, *QC = f + L.QC, *W = f + L.W, *dQC = f + L.dQC, *dS0 = f + L.dS0, *dS1 = f + L.dS1, *dc0 = f + L.dc0,
          *dc1 = f + L.dc1, *dh1 = f + L.dh1, *DE = f + L.DE, *dGemb = f + L.dGemb, *tmpq = f + L.tmpq, *csw = f + L.cs_scratch,
          *W2 = f + L.W2, *FM = f + L.FM;
    const int* tok = s->iws + L.tok;
    const size_t fsz = sizeof(float);
    const long long SB = (long long)S * B;
    const bool tc = s->use_tc != 0;
    __nv_bfloat16 *Wcat0b = (__nv_bfloat16*)(f + L.Wcat0b), *Wcat1b = (__nv_bfloat16*)(f + L.Wcat1b), *Wqb = (__nv_bfloat16*)(f + L.Wqb),
                  *S0b = (__nv_bfloat16*)(f + L.S0b), *S1b = (__nv_bfloat16*)(f + L.S1b), *G0b = (__nv_bfloat16*)(f + L.G0b),
                  *G1b = (__nv_bfloat16*)(f + L.G1b), *dQb = (__nv_bfloat16*)(f + L.dQb);

    const bool loop_phase = (phases & 1) != 0, param_phase = (phases & 2) != 0;
    // dQC[1..S] = dlogits . emb  (row m = t*B + b  <-  dlogits[b, t, :])
    if (loop_phase) {
    LAS_CUDA(cudaMemsetAsync(dQC, 0, (size_t)B * 2 * P * fsz, st));
        LasGemmF32 d{};
        d.A = g->dlogits; d.B = s->emb; d.C = dQC + (size_t)B * 2 * P;
        d.M = (int)SB; d.N = 2 * P; d.K = V; d.batch = 1;
        d.a_m_inner = B; d.a_m_so = V; d.a_m_si = (long long)S * V; d.a_k_si = 1;
        d.b_k_si = E; d.b_n_s = 1; d.c_m_si = 2 * P;
        d.alpha = 1.f; d.beta = 0.f;
        RC(las_gemm_f32(&d, st));
    }
    __nv_bfloat16 *dlb = (__nv_bfloat16*)(f + L.dlb), *ohb = (__nv_bfloat16*)(f + L.ohb), *QCb = (__nv_bfloat16*)(f + L.QCb);
    float *tmp32 = f + L.tmp32, *skws = f + L.skws;
    const bool tc_tok = tc && V <= 32;
    // tied classifier weight: d_emb = dlogits^T . QC[1..S]  (parameter-gradient phase; reads only dlogits and the forward's QC rows,
    // every scratch buffer below has its own region of the workspace, so it does not matter whether it runs before or after the loop)
    if (param_phase) {
    if (tc_tok) {
        // (S*B, 32) bf16 zero-padded dlogits in (t, b) row order and bf16 QC rows -> one split-K tensor-core GEMM
        RC(las_cast_f32_to_bf16(g->dlogits, (long long)S * V, B, V, dlb, 32, SB, V, 32, st));
        RC(cast_rows(st, QC + (size_t)B * 2 * P, 2 * P, QCb, 2 * P, SB, 2 * P));
        RC(tc_tn(st, dlb, 32, QCb, 2 * P, tmp32, 2 * P, 32, 2 * P, (int)SB, skws, L.skws_floats));
        LAS_CUDA(cudaMemcpyAsync(g->d_emb, tmp32, (size_t)V * E * fsz, cudaMemcpyDeviceToDevice, st));
    } else {
        LasGemmF32 d{};
        d.A = g->dlogits; d.B = QC + (size_t)B * 2 * P; d.C = g->d_emb;
        d.M = V; d.N = 2 * P; d.K = (int)SB; d.batch = 1;
        d.a_m_si = 1; d.a_k_inner = B; d.a_k_so = V; d.a_k_si = (long long)S * V;
        d.b_k_si = 2 * P; d.b_n_s = 1; d.c_m_si = E;
        d.alpha = 1.f; d.beta = 0.f;
        RC(las_gemm_f32(&d, st));
    }
    RC(las_colsum_f32(g->dlogits, V, (int)SB, V, g->d_cls_b, 0, csw, st));
    }

    LasAttnStep at{};
    at.K = s->K; at.V = s->V_; at.lens = s->enc_lens; at.B = B; at.T = T; at.P = P; at.heads = heads;
    at.scale = sqrtf((float)d_head);
    at.kv_bf16 = s->kv_bf16;
    at.ld_q = 2 * P; at.ld_w = T; at.ld_dctx = 2 * P; at.ld_dctx2 = K0; at.ld_dq = 2 * P; at.dq_accumulate = 1;
    PlanBuf bq1, bq2, bq3;
    if (tc) {
        LAS_CHECK_ARG(las_tc_plan_bytes() <= sizeof(PlanBuf), "speller: plan buffer too small");
        RC(las_tc_plan_make(&bq1, dQb, Wqb, B, DO, P, S + 1, P, (long long)B * P, DO, 1));
        RC(las_tc_plan_make(&bq2, G1b, Wcat1b, B, K1, 4 * DO, S, 4 * DO, (long long)B * 4 * DO, K1, 1));
        RC(las_tc_plan_make(&bq3, G0b, Wcat0b, B, K0, 4 * DH, S, 4 * DH, (long long)B * 4 * DH, K0, 1));
    }
    // dS1 / dS0 have 12 output tiles and 16 / 32 K iterations: split K over the idle SMs and leave the partials un-reduced;
    // cell_bwd_kernel and the attention backward add them up as they read
    const int skb1 = tc ? pick_split(las_tc_plan_tiles(&bq2), las_tc_plan_kiters(&bq2)) : 1;
    const int skb0 = tc ? pick_split(las_tc_plan_tiles(&bq3), las_tc_plan_kiters(&bq3)) : 1;
    if (tc) { dS1 = f + L.dSp1; dS0 = f + L.dSp0; }
    const long long st1 = (long long)B * K1, st0 = (long long)B * K0;

    if (loop_phase) {
    {
    LasPdlScope pdl_scope;       // the per-step kernels overlap their launch / prologue with the predecessor's tail
    for (int t = S - 1; t >= 0; --t) {
        RC(seg_step(seg, S - 1 - t));
        const int rn = t + 1;
        float* dQCn = dQC + (size_t)rn * B * 2 * P;
        // attention step (t+1): dctx_total = classifier path + cell-0 path of step t+1
        at.q = QC + (size_t)rn * B * 2 * P; at.w = W + (size_t)rn * B * heads * T;
        at.ctx = QC + (size_t)rn * B * 2 * P + P; at.ld_ctx = 2 * P;        // saved context: sum_t w_t (dctx.V_t) == dctx.ctx
        at.dctx = dQCn + P; at.dctx2 = (t == S - 1) ? nullptr : dS0;
        at.dctx2_nsplit = skb0; at.dctx2_split_stride = st0;
        at.dq = dQCn; at.de = DE + (size_t)rn * B * heads * T;
        at.dq_bf16 = tc ? (void*)(dQb + (size_t)rn * B * P) : nullptr; at.ld_dq_bf16 = P;
        if (s->init_force) { at.fmask = FM + (size_t)t * T; at.ld_fmask = 0; at.w2 = W2 + (size_t)rn * B * heads * T; }
        const bool fuse_tail = tc && las_attn_step_bwd_cell_supported(&at, DO) != 0;      // asked per step: the descriptor is complete here
        CellBwd b1{};
        b1.G = G1 + (size_t)t * B * 4 * DO;
        b1.dh_a = dh1; b1.ld_a = DO;
        b1.dh_b = (t == S - 1) ? nullptr : dS1 + DH; b1.ld_b = K1; b1.nsplit_b = skb1; b1.stride_b = st1;
        b1.mask = s->drop1 ? s->drop1 + (size_t)t * B * DO : nullptr;
        b1.c = C1 + (size_t)rn * B * DO; b1.ld_c = DO; b1.c_prev = C1 + (size_t)t * B * DO; b1.ld_cp = DO;
        b1.dc = dc1; b1.first = (t == S - 1); b1.B = B; b1.H = DO;
        b1.Gb = tc ? G1b + (size_t)t * B * 4 * DO : nullptr;
        if (fuse_tail) {
            // attention backward + dh1 = dq_total . Wq + cell-1 backward in ONE launch (attn_tail.h)
            LasAttnCellTail tl{};
            tl.wq_bf16 = Wqb; tl.DO = DO; tl.G = b1.G; tl.dh_b = b1.dh_b; tl.ld_b = b1.ld_b; tl.stride_b = b1.stride_b; tl.nsplit_b = b1.nsplit_b;
            tl.mask = b1.mask; tl.c = b1.c; tl.ld_c = b1.ld_c; tl.c_prev = b1.c_prev; tl.ld_cp = b1.ld_cp; tl.dc = b1.dc; tl.Gb = b1.Gb;
            tl.first = b1.first;
            tl.K_f16 = s->K_f16; tl.V_f16 = s->V_f16;
            RC(las_attn_step_bwd_cell(&at, &tl, st));
        } else {
            RC(las_attn_step_bwd_f32(&at, st));
            // dh1_t (dropped) = dq_total . Wq  (+ recurrent path, added inside cell_bwd)
            if (tc) RC(las_tc_plan_launch(&bq1, rn, dh1, DO, nullptr, nullptr, st));
            else RC(gemm(st, dQCn, 2 * P, s->wq, DO, 0, dh1, DO, B, DO, P));
            LAS_CUDA(las_launch(cell_bwd_kernel, dim3(ceil_div(B * DO, 256)), dim3(256), 0, st, b1));
            LAS_LAUNCH_CHECK();
        }
        // dS1[t] = dG1_t . Wcat1  -> [dh0_t | dh1_{t-1}]
        if (tc) RC(las_tc_plan_launch_split(&bq2, t, dS1, K1, skb1, st));
        else RC(gemm(st, b1.G, 4 * DO, Wcat1, K1, 0, dS1, K1, B, K1, 4 * DO));
        CellBwd b0{};
        b0.G = G0 + (size_t)t * B * 4 * DH;
        b0.dh_a = dS1; b0.ld_a = K1; b0.nsplit_a = skb1; b0.stride_a = st1;
        b0.dh_b = (t == S - 1) ? nullptr : dS0 + P; b0.ld_b = K0; b0.nsplit_b = skb0; b0.stride_b = st0;
        b0.mask = s->drop0 ? s->drop0 + (size_t)t * B * DH : nullptr;
        b0.c = C0 + (size_t)rn * B * DH; b0.ld_c = DH; b0.c_prev = C0 + (size_t)t * B * DH; b0.ld_cp = DH;
        b0.dc = dc0; b0.first = (t == S - 1); b0.B = B; b0.H = DH;
        b0.Gb = tc ? G0b + (size_t)t * B * 4 * DH : nullptr;
        LAS_CUDA(las_launch(cell_bwd_kernel, dim3(ceil_div(B * DH, 256)), dim3(256), 0, st, b0));
        LAS_LAUNCH_CHECK();
        // dS0[t] = dG0_t . Wcat0 -> [dctx_t | dh0_{t-1}]
        if (tc) RC(las_tc_plan_launch_split(&bq3, t, dS0, K0, skb0, st));
        else RC(gemm(st, b0.G, 4 * DH, Wcat0, K0, 0, dS0, K0, B, K0, 4 * DH));
    }
    }                            // PDL scope ends: the batched GEMMs below are ordinary launches
    // initial attention (src/models.py:346): its context feeds cell 0 of step 0 only
    at.q = QC; at.w = W; at.ctx = QC + P; at.dctx = dQC + P; at.dctx2 = dS0; at.dq = dQC; at.de = DE;
    at.dq_bf16 = tc ? (void*)dQb : nullptr;
    at.fmask = nullptr; at.w2 = nullptr;
    RC(las_attn_step_bwd_f32(&at, st));
    // keys / values: dK[b] = DE[:, b]^T . Q[:, b] ; dV[b] = W[:, b]^T . dctx[:, b]   (batched over b, per head)
    for (int h = 0; h < heads; ++h) {
        LasGemmF32 d{};
        d.M = T; d.N = d_head; d.K = S + 1; d.batch = B;
        d.a_m_si = 1; d.a_k_si = (long long)B * heads * T; d.bsA = (long long)heads * T;
        d.b_k_si = (long long)B * 2 * P; d.b_n_s = 1; d.bsB = 2 * P;
        d.c_m_si = P; d.bsC = (long long)T * P;
        d.alpha = 1.f; d.beta = 0.f;
        d.A = DE + (size_t)h * T; d.B = QC + (size_t)h * d_head; d.C = g->dK + (size_t)h * d_head;
        RC(las_gemm_f32(&d, st));
        d.A = (s->init_force ? W2 : W) + (size_t)h * T; d.B = dQC + P + (size_t)h * d_head; d.C = g->dV + (size_t)h * d_head;
        RC(las_gemm_f32(&d, st));
    }
    }   // loop_phase

    // ---- batched parameter gradients (phase 2: they read what the loop left in the workspace and feed only the optimizer) ----
    if (param_phase) {
    // query_map: rows 1..S see h1_t (S1[t+1] slot), row 0 sees init_query
    if (tc) RC(tc_tn(st, dQb + (size_t)B * P, P, S1b + (size_t)B * K1 + DH, K1, g->d_wq, DO, P, DO, (int)SB, skws, L.skws_floats));
    else RC(gemm_tn(st, dQC + (size_t)B * 2 * P, 2 * P, S1 + (size_t)B * K1 + DH, K1, g->d_wq, DO, P, DO, (int)SB));
    {
        LasGemmF32 d{};
        d.A = dQC; d.B = s->init_query; d.C = g->d_wq;
        d.M = P; d.N = DO; d.K = B; d.batch = 1;
        d.a_m_si = 1; d.a_k_si = 2 * P; d.b_k_si = 0; d.b_n_s = 1; d.c_m_si = DO;
        d.alpha = 1.f; d.beta = 1.f;
        RC(las_gemm_f32(&d, st));
    }
    RC(las_colsum_f32(dQC, 2 * P, (S + 1) * B, P, g->d_bq, 0, csw, st));
    RC(gemm(st, dQC, 2 * P, s->wq, DO, 0, tmpq, DO, B, DO, P));
    RC(las_colsum_f32(tmpq, DO, B, DO, g->d_init_query, 0, csw, st));
    // cell 1
    if (tc) {
        RC(tc_tn(st, G1b, 4 * DO, S1b, K1, g->d_w_ih1, DH, 4 * DO, DH, (int)SB, skws, L.skws_floats));
        RC(tc_tn(st, G1b, 4 * DO, S1b + DH, K1, g->d_w_hh1, DO, 4 * DO, DO, (int)SB, skws, L.skws_floats));
    } else {
        RC(gemm_tn(st, G1, 4 * DO, S1, K1, g->d_w_ih1, DH, 4 * DO, DH, (int)SB));
        RC(gemm_tn(st, G1, 4 * DO, S1 + DH, K1, g->d_w_hh1, DO, 4 * DO, DO, (int)SB));
    }
    RC(las_colsum_f32(G1, 4 * DO, (int)SB, 4 * DO, g->d_b_ih1, 0, csw, st));
    LAS_CUDA(cudaMemcpyAsync(g->d_b_hh1, g->d_b_ih1, (size_t)4 * DO * fsz, cudaMemcpyDeviceToDevice, st));
    // cell 0: context columns, recurrent weight, biases
    if (tc) {
        RC(tc_tn(st, G0b, 4 * DH, S0b, K0, g->d_w_ih0 + E, E + P, 4 * DH, P, (int)SB, skws, L.skws_floats));
        RC(tc_tn(st, G0b, 4 * DH, S0b + P, K0, g->d_w_hh0, DH, 4 * DH, DH, (int)SB, skws, L.skws_floats));
    } else {
        RC(gemm_tn(st, G0, 4 * DH, S0, K0, g->d_w_ih0 + E, E + P, 4 * DH, P, (int)SB));
        RC(gemm_tn(st, G0, 4 * DH, S0 + P, K0, g->d_w_hh0, DH, 4 * DH, DH, (int)SB));
    }
    RC(las_colsum_f32(G0, 4 * DH, (int)SB, 4 * DH, g->d_b_ih0, 0, csw, st));
    LAS_CUDA(cudaMemcpyAsync(g->d_b_hh0, g->d_b_ih0, (size_t)4 * DH * fsz, cudaMemcpyDeviceToDevice, st));
    // cell 0: embedding columns through the token table
    if (tc_tok) {
        // dGemb = OneHot(tok)^T . dG0 : the token-wise scatter-reduce as a split-K tensor-core GEMM
        onehot_kernel<<<ceil_div((int)SB * 32, 256), 256, 0, st>>>(tok, ohb, (int)SB);
        LAS_LAUNCH_CHECK();
        RC(tc_tn(st, ohb, 32, G0b, 4 * DH, tmp32, 4 * DH, 32, 4 * DH, (int)SB, skws, L.skws_floats));
        LAS_CUDA(cudaMemcpyAsync(dGemb, tmp32, (size_t)V * 4 * DH * fsz, cudaMemcpyDeviceToDevice, st));
    } else {
        token_reduce_kernel<<<dim3(ceil_div(4 * DH, 256), V), 256, 0, st>>>(G0, tok, dGemb, (int)SB, 4 * DH);
        LAS_LAUNCH_CHECK();
    }
    RC(gemm_tn(st, dGemb, 4 * DH, s->emb, E, g->d_w_ih0, E + P, 4 * DH, E, V));
    // embedding rows via lookups: padding_idx row receives no lookup gradient (nn.Embedding(padding_idx), src/models.py:261-265)
    if (s->pad_idx >= 0 && s->pad_idx < V) LAS_CUDA(cudaMemsetAsync(dGemb + (size_t)s->pad_idx * 4 * DH, 0, (size_t)4 * DH * fsz, st));
    if (tc && (4 * DH) % 16 == 0 && (size_t)(16 + CS_ROWSPLIT) * V * E <= L.skws_floats) {
        // d_emb += dGemb . W_ih0[:, :E] has M = V (30) rows, i.e. 16 output tiles for a K = 4*DH reduction (0.3 ms on 16 CTAs): cut K
        // into 16 batched chunks (256 CTAs) and let a column-sum pass add the partial products into d_emb
        const int NCH = 16, kc = 4 * DH / NCH;
        LasGemmF32 d{};
        d.A = dGemb; d.B = s->w_ih0; d.C = skws;
        d.M = V; d.N = E; d.K = kc; d.batch = NCH;
        d.a_m_si = 4 * DH; d.a_k_si = 1; d.bsA = kc;
        d.b_k_si = E + P; d.b_n_s = 1; d.bsB = (long long)kc * (E + P);
        d.c_m_si = E; d.bsC = (long long)V * E;
        d.alpha = 1.f; d.beta = 0.f;
        RC(las_gemm_f32(&d, st));
        RC(las_colsum_f32(skws, (long long)V * E, NCH, V * E, g->d_emb, 1, skws + (size_t)NCH * V * E, st));
    } else {
        RC(gemm(st, dGemb, 4 * DH, s->w_ih0, E + P, 0, g->d_emb, E, V, E, 4 * DH, 1.f));
    }
    }   // param_phase
    return LAS_OK;
}
}  // namespace
