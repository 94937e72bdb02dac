// Fused single-query attention step: energy -> x sqrt(d) -> length mask -> softmax -> zero pads -> context.
//
// Replaces MultiheadCrossAttention.forward after the query projection (reference src/models.py:168-185):
//   e = (q . K^T) / norm_factor  with norm_factor = 1/sqrt(d)  => e = q . K^T * sqrt(d)   (:93, :170)
//   e[t >= len] = finfo.min ; w = softmax(e) ; w[t >= len] = 0 ; ctx = w . V                (:171-185)
// One CTA per (batch row, head).  Keys and values are (B, T_enc, P) row-major: each warp streams whole rows with
// 128-bit loads (lane l reads floats [4l, 4l+4) and [128+4l, ...)), reduces the dot product with warp shuffles,
// and K and V are each read exactly once per step: algorithmic bytes = 2 * B * T_enc * P * 4.
#include "las_common.cuh"
#include "las_b200.h"
#include "attn_tail.h"
#include <float.h>
#include <stdlib.h>

namespace {

constexpr int NT = 512;     // 16 warps: 4 rows in flight per warp -> ~64 KB of loads in flight per SM
constexpr int NW = NT / 32;
constexpr int MAXCH = 2;     // head dim <= 256 (2 x 128-float chunks per warp row pass)

// K / V rows: plain read-only loads.  ld.global.nc.L1::no_allocate was measured at HALF the per-SM streaming rate of the allocating
// form on B200 (persistent decoder kernel, 410 KB per SM per step: 6.0 us vs 3.3 us per pass), so the "streaming" hint is not used.
__device__ __forceinline__ float4 ldg4_stream(const float* p) {
    return __ldg(reinterpret_cast<const float4*>(p));
}

// 4 consecutive elements of a K/V row as floats: fp32 rows -> one 128-bit load; bf16 rows (AMP mode) -> one 64-bit load
template <bool KV16>
__device__ __forceinline__ float4 load_kv4(const void* base, long long elem) {
    if (KV16) {
        uint2 r;
        asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"((const __nv_bfloat16*)base + elem));
        return make_float4(__uint_as_float(r.x << 16), __uint_as_float(r.x & 0xffff0000u), __uint_as_float(r.y << 16),
                           __uint_as_float(r.y & 0xffff0000u));
    }
    return ldg4_stream((const float*)base + elem);
}

__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
    v = is_max ? warp_max(v) : warp_sum(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    float r = red[0];
#pragma unroll
    for (int i = 1; i < NW; ++i) r = is_max ? fmaxf(r, red[i]) : (r + red[i]);
    return r;
}

// phase A: s[t] = scale * dot(vec, M[t]) for t < len, over rows of M (B,T,P) restricted to one head
// phase C: out[p] = sum_t s2[t] * M2[t][p]
// Used as fwd (vec=q, M=K, M2=V) and bwd (vec=dctx, M=V, M2=K).
template <bool BWD, int RU, bool KV16>
__global__ void __launch_bounds__(NT) attn_step_kernel(LasAttnStep a) {
    extern __shared__ __align__(16) float sm[];
    pdl_wait();
    pdl_trigger();
    const int T = a.T, P = a.P, heads = a.heads, d = P / heads;
    const int Tp = (T + 3) & ~3;    // keep `part` 16-byte aligned
    float* sc = sm;                 // [Tp]  scores / weights
    float* red = sm + Tp;           // [NW]
    float* part = red + NW;         // [NW][d] cross-warp partials for phase C
    const int bh = blockIdx.x, b = bh / heads, h = bh - b * heads;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int len = min(a.lens[b], T);
    const void* Mat1 = BWD ? (const void*)a.V : (const void*)a.K;      // fp32 or bf16 (KV16) elements
    const void* Mat2 = BWD ? (const void*)a.K : (const void*)a.V;
    const long long mbase = (long long)b * T * P + h * d;
    const float* vec = (BWD ? a.dctx + (long long)b * a.ld_dctx : a.q + (long long)b * a.ld_q) + h * d;
    const int nch = (d + 127) / 128;

    float4 v4[MAXCH];
#pragma unroll
    for (int c = 0; c < MAXCH; ++c) {
        int k = c * 128 + lane * 4;
        v4[c] = (c < nch && k < d) ? *reinterpret_cast<const float4*>(vec + k) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (BWD && a.dctx2 && c < nch && k < d) {
            // dctx_total = dctx (classifier path) + dctx2 (next step's cell-0 input path); the sum is written back so
            // the deferred dV = w^T . dctx GEMM sees it.  Every warp computes the same sum; warp 0 stores it.
            // split-K partials of the producing GEMM (at most 8): all loads issued before the adds
            float4 e[8];
            const int ns = a.dctx2_nsplit > 1 ? a.dctx2_nsplit : 1;
#pragma unroll
            for (int sp = 0; sp < 8; ++sp)
                e[sp] = sp < ns ? *reinterpret_cast<const float4*>(a.dctx2 + sp * a.dctx2_split_stride + (long long)b * a.ld_dctx2 + h * d + k)
                                : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int sp = 0; sp < 8; ++sp) { v4[c].x += e[sp].x; v4[c].y += e[sp].y; v4[c].z += e[sp].z; v4[c].w += e[sp].w; }
        }
    }
    if (BWD && a.dctx2) {
        __syncthreads();   // all warps have read dctx before warp 0 overwrites it
        if (w == 0) {
#pragma unroll
            for (int c = 0; c < MAXCH; ++c) {
                int k = c * 128 + lane * 4;
                if (c < nch && k < d) *reinterpret_cast<float4*>(a.dctx + (long long)b * a.ld_dctx + h * d + k) = v4[c];
            }
        }
    }
    // init-force prior (src/models.py:177-181): the second softmax runs over ALL T positions (pads included), so the value
    // pass of forward and backward covers T rows instead of len
    const float* fm = a.fmask ? a.fmask + ((long long)b * heads + h) * a.ld_fmask : nullptr;
    const int lenA = (BWD && fm) ? T : len, lenC = (!BWD && fm) ? T : len;
    // ---- phase A: one warp per row, RU rows in flight per warp (all loads issued before any reduction) ----
    for (int t0 = w; t0 < lenA; t0 += NW * RU) {
        float4 m[RU][MAXCH];
#pragma unroll
        for (int r = 0; r < RU; ++r) {
            const int t = t0 + r * NW;
#pragma unroll
            for (int c = 0; c < MAXCH; ++c) {
                int k = c * 128 + lane * 4;
                m[r][c] = (t < lenA && c < nch && k < d) ? load_kv4<KV16>(Mat1, mbase + (long long)t * P + k) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll
        for (int r = 0; r < RU; ++r) {
            const int t = t0 + r * NW;
            float acc = 0.f;
#pragma unroll
            for (int c = 0; c < MAXCH; ++c) {
                acc = fmaf(m[r][c].x, v4[c].x, acc); acc = fmaf(m[r][c].y, v4[c].y, acc);
                acc = fmaf(m[r][c].z, v4[c].z, acc); acc = fmaf(m[r][c].w, v4[c].w, acc);
            }
            acc = warp_sum(acc);
            if (lane == 0 && t < lenA) sc[t] = acc;
        }
    }
    __syncthreads();
    float* wrow = a.w + ((long long)b * heads + h) * a.ld_w;
    float* w2row = fm ? a.w2 + ((long long)b * heads + h) * a.ld_w : nullptr;
    if (!BWD) {
        // ---- softmax over t < len (masked entries: exp(finfo.min - max) == 0 exactly, then forced to 0) ----
        float mx = -FLT_MAX;
        for (int t = tid; t < len; t += NT) mx = fmaxf(mx, sc[t] * a.scale);
        mx = block_reduce(mx, red, true);
        float sum = 0.f;
        for (int t = tid; t < len; t += NT) {
            float e = expf(sc[t] * a.scale - mx);
            sc[t] = e;
            sum += e;
        }
        sum = block_reduce(sum, red, false);
        const float inv = 1.f / sum;
        for (int t = tid; t < T; t += NT) {
            float wv = (t < len) ? sc[t] * inv : 0.f;
            if (t < len) sc[t] = wv;
            wrow[t] = wv;
            if (b == 0 && a.w_b0) a.w_b0[(long long)h * T + t] = wv;
        }
        if (fm) {
            // w2 = softmax(w * prior) over all T positions; the returned / recorded weights stay the pre-prior ones (:178,188)
            float mx2 = -FLT_MAX;
            for (int t = tid; t < T; t += NT) {
                const float x = (t < len) ? sc[t] * fm[t] : 0.f;
                sc[t] = x;
                mx2 = fmaxf(mx2, x);
            }
            mx2 = block_reduce(mx2, red, true);
            float sum2 = 0.f;
            for (int t = tid; t < T; t += NT) {
                const float e = expf(sc[t] - mx2);
                sc[t] = e;
                sum2 += e;
            }
            sum2 = block_reduce(sum2, red, false);
            const float inv2 = 1.f / sum2;
            for (int t = tid; t < T; t += NT) {
                const float wv = sc[t] * inv2;
                sc[t] = wv;
                w2row[t] = wv;
            }
        }
    } else {
        if (fm) {
            // sc[t] = dw2[t] for all t < T -> d(w * prior) = w2 (dw2 - sum w2 dw2) -> dw[t] = that * prior[t] (0 at pads)
            float dot2 = 0.f;
            for (int t = tid; t < T; t += NT) dot2 = fmaf(w2row[t], sc[t], dot2);
            dot2 = block_reduce(dot2, red, false);
            for (int t = tid; t < T; t += NT) sc[t] = (t < len) ? w2row[t] * (sc[t] - dot2) * fm[t] : 0.f;
        }
        // dw[t] = sc[t]; de[t] = w[t] * (dw[t] - sum_t' w[t'] dw[t']) ; stored pre-multiplied by scale
        float dot = 0.f;
        for (int t = tid; t < len; t += NT) dot = fmaf(wrow[t], sc[t], dot);
        dot = block_reduce(dot, red, false);
        float* derow = a.de + ((long long)b * heads + h) * a.ld_w;
        for (int t = tid; t < T; t += NT) {
            float de = (t < len) ? wrow[t] * (sc[t] - dot) * a.scale : 0.f;
            if (t < len) sc[t] = de;
            derow[t] = de;
        }
    }
    __syncthreads();
    // ---- phase C: out[p] = sum_t sc[t] * Mat2[t][p] ----
    float4 o4[MAXCH];
#pragma unroll
    for (int c = 0; c < MAXCH; ++c) o4[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int t0 = w; t0 < lenC; t0 += NW * RU) {
        float4 m[RU][MAXCH];
        float sv[RU];
#pragma unroll
        for (int r = 0; r < RU; ++r) {
            const int t = t0 + r * NW;
            sv[r] = (t < lenC) ? sc[t] : 0.f;
#pragma unroll
            for (int c = 0; c < MAXCH; ++c) {
                int k = c * 128 + lane * 4;
                m[r][c] = (t < lenC && c < nch && k < d) ? load_kv4<KV16>(Mat2, mbase + (long long)t * P + k) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll
        for (int r = 0; r < RU; ++r)
#pragma unroll
            for (int c = 0; c < MAXCH; ++c) {
                o4[c].x = fmaf(sv[r], m[r][c].x, o4[c].x); o4[c].y = fmaf(sv[r], m[r][c].y, o4[c].y);
                o4[c].z = fmaf(sv[r], m[r][c].z, o4[c].z); o4[c].w = fmaf(sv[r], m[r][c].w, o4[c].w);
            }
    }
#pragma unroll
    for (int c = 0; c < MAXCH; ++c) {
        int k = c * 128 + lane * 4;
        if (c < nch && k < d) *reinterpret_cast<float4*>(part + w * d + k) = o4[c];
    }
    __syncthreads();
    for (int p = tid; p < d; p += NT) {
        float r = 0.f;
#pragma unroll
        for (int i = 0; i < NW; ++i) r += part[i * d + p];
        if (!BWD) {
            a.ctx[(long long)b * a.ld_ctx + h * d + p] = r;
            if (a.ctx2) a.ctx2[(long long)b * a.ld_ctx2 + h * d + p] = r;
            if (a.ctx2_bf16) ((__nv_bfloat16*)a.ctx2_bf16)[(long long)b * a.ld_ctx2_bf16 + h * d + p] = __float2bfloat16(r);
        } else {
            float* dq = a.dq + (long long)b * a.ld_dq + h * d + p;
            const float tot = (a.dq_accumulate ? *dq : 0.f) + r;
            *dq = tot;
            if (a.dq_bf16) ((__nv_bfloat16*)a.dq_bf16)[(long long)b * a.ld_dq_bf16 + h * d + p] = __float2bfloat16(tot);
        }
    }
}


// ---------------------------------------------------------------------------------------------------------------------
// Single-pass, T-split variant (the default path).
//   * grid (S, B*heads), cluster (S,1,1): the S CTAs of a cluster split the valid rows [0, len) of one (batch row, head),
//     so B = 96 fills all 148 SMs instead of 96 of them;
//   * K[t] and V[t] of a row are loaded TOGETHER (all loads of an iteration are issued before any arithmetic), so the
//     kernel is one pass over memory with no dependent second phase:
//       fwd: online softmax (running max / sum / weighted V sum per warp), merged across warps in shared memory and
//            across the cluster through distributed shared memory; w[t] = exp(e_t - M) / S is written once M, S are final;
//       bwd: sum_t w_t (dctx . V_t) == dctx . ctx, so with the saved context the softmax backward needs no pre-pass:
//            de_t = w_t (dctx . V_t - dctx . ctx) scale ;  dq = sum_t de_t K_t.
// K and V are still read exactly once per step.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int NT2 = 256;
constexpr int NW2 = NT2 / 32;
constexpr int RU2 = 4;        // (K row + V row) x 4 in flight per warp: 8 KB (fp32, d = 256)

__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float ld_dsmem_f32(const float* local_ptr, unsigned rank) {
    unsigned la = (unsigned)__cvta_generic_to_shared(local_ptr), ra;
    float v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(rank));
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(ra) : "memory");
    return v;
}

constexpr int TAIL_MAX_SPLIT = 8;

template <bool BWD, bool KV16, bool TAIL>
__global__ void __launch_bounds__(NT2, 2) attn_step_split_kernel(LasAttnStep a, LasAttnCellTail tl) {
    extern __shared__ __align__(16) float sm[];
    const int T = a.T, P = a.P, heads = a.heads, d = P / heads;
    const int S = gridDim.x, rank = blockIdx.x;
    const int bh = blockIdx.y, b = bh / heads, h = bh - b * heads;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int cap = ((T + S - 1) / S + 3) & ~3;
    // fused tail (attn_tail.h): this CTA's DO/S output columns of dh1 = dq . Wq, then LSTMCell-1 backward for them
    const int nper = TAIL ? tl.DO / S : 8;
    float* dqs = sm + cap + 2 * NW2 + 16 + 4 + d + NW2 * d;       // [d]          dq_total of the row (every CTA of the cluster holds all of it)
    float* gpart = dqs + d;                                       // [kg][nper]   partial dh1 sums of the k groups
    __nv_bfloat16* wq_s = reinterpret_cast<__nv_bfloat16*>(gpart + (TAIL ? (NT2 / (nper / 2)) * nper : 0));      // [P][nper]
    if constexpr (TAIL) {
        // the weight slice does not depend on the predecessor kernel: fetch it before the programmatic-launch wait, asynchronously
        const int cpr = nper / 8;                                 // 16-byte chunks per weight row
        const __nv_bfloat16* wsrc = reinterpret_cast<const __nv_bfloat16*>(tl.wq_bf16) + rank * nper;
        for (int i = tid; i < P * cpr; i += NT2) {
            const int k = i / cpr, j = i - k * cpr;
            cp_async16(wq_s + k * nper + j * 8, wsrc + (long long)k * tl.DO + j * 8);
        }
        cp_async_commit();
    }
    pdl_wait();
    pdl_trigger();
    const int len = min(a.lens[b], T);
    const int per = (len + S - 1) / S;                       // rows of this (row, head) per CTA
    const int t_lo = min(rank * per, len), t_hi = min(t_lo + per, len);
    float* sc = sm;                 // [cap]    fwd: scaled energies of this CTA's rows
    float* wm = sm + cap;           // [NW2]    per-warp running max
    float* wsum = wm + NW2;         // [NW2]    per-warp running sum
    float* cl = wsum + NW2;         // [2*8]    (m, s) of every CTA of the cluster, gathered
    float* cta_ms = cl + 16;        // [4]      this CTA's (m, s)
    float* cta_o = cta_ms + 4;      // [d]      this CTA's partial output vector
    float* part = cta_o + d;        // [NW2][d] per-warp partial output vectors
    const void* MatK = a.K;
    const void* MatV = a.V;
    const long long mbase = (long long)b * T * P + h * d;
    const float* vec = (BWD ? a.dctx + (long long)b * a.ld_dctx : a.q + (long long)b * a.ld_q) + h * d;
    const int nch = (d + 127) / 128;

    float4 v4[MAXCH];
    float dot = 0.f;
#pragma unroll
    for (int c = 0; c < MAXCH; ++c) {
        const int k = c * 128 + lane * 4;
        const bool in = c < nch && k < d;
        v4[c] = in ? *reinterpret_cast<const float4*>(vec + k) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (BWD && in) {
            if (a.dctx2) {      // dctx_total = classifier path + next step's cell-0 input path (summed over its split-K partials)
                float4 e[8];                      // at most 8 partials: all loads issued before the adds
                const int ns = a.dctx2_nsplit > 1 ? a.dctx2_nsplit : 1;
#pragma unroll
                for (int sp = 0; sp < 8; ++sp)
                    e[sp] = sp < ns ? *reinterpret_cast<const float4*>(a.dctx2 + sp * a.dctx2_split_stride + (long long)b * a.ld_dctx2 + h * d + k)
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int sp = 0; sp < 8; ++sp) { v4[c].x += e[sp].x; v4[c].y += e[sp].y; v4[c].z += e[sp].z; v4[c].w += e[sp].w; }
            }
            const float4 c4 = *reinterpret_cast<const float4*>(a.ctx + (long long)b * a.ld_ctx + h * d + k);
            dot = fmaf(v4[c].x, c4.x, dot); dot = fmaf(v4[c].y, c4.y, dot); dot = fmaf(v4[c].z, c4.z, dot); dot = fmaf(v4[c].w, c4.w, dot);
        }
    }
    if (BWD) dot = warp_sum(dot);
    const float* wrow_c = a.w + ((long long)b * heads + h) * a.ld_w;

    float m_run = -INFINITY, s_run = 0.f;
    float4 o4[MAXCH];
#pragma unroll
    for (int c = 0; c < MAXCH; ++c) o4[c] = make_float4(0.f, 0.f, 0.f, 0.f);

    for (int tb = t_lo + w; tb < t_hi; tb += NW2 * RU2) {
        float4 kk[RU2][MAXCH], vv[RU2][MAXCH];
        float wt[RU2];
#pragma unroll
        for (int r = 0; r < RU2; ++r) {
            const int t = tb + r * NW2;
            const bool ok = t < t_hi;
#pragma unroll
            for (int c = 0; c < MAXCH; ++c) {
                const int k = c * 128 + lane * 4;
                const bool in = ok && c < nch && k < d;
                kk[r][c] = in ? load_kv4<KV16>(MatK, mbase + (long long)t * P + k) : make_float4(0.f, 0.f, 0.f, 0.f);
                vv[r][c] = in ? load_kv4<KV16>(MatV, mbase + (long long)t * P + k) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            wt[r] = (BWD && ok) ? wrow_c[t] : 0.f;
        }
        float e[RU2];
#pragma unroll
        for (int r = 0; r < RU2; ++r) {
            float acc = 0.f;
#pragma unroll
            for (int c = 0; c < MAXCH; ++c) {
                const float4 m = BWD ? vv[r][c] : kk[r][c];
                acc = fmaf(m.x, v4[c].x, acc); acc = fmaf(m.y, v4[c].y, acc);
                acc = fmaf(m.z, v4[c].z, acc); acc = fmaf(m.w, v4[c].w, acc);
            }
            e[r] = warp_sum(acc);
        }
        if (!BWD) {
            float m_new = m_run;
#pragma unroll
            for (int r = 0; r < RU2; ++r) {
                const int t = tb + r * NW2;
                e[r] = (t < t_hi) ? e[r] * a.scale : -INFINITY;
                if (lane == 0 && t < t_hi) sc[t - t_lo] = e[r];
                m_new = fmaxf(m_new, e[r]);
            }
            const float f = expf(m_run - m_new);          // first iteration: exp(-inf) = 0 ; row tb itself is valid, so m_new is finite
            s_run *= f;
#pragma unroll
            for (int c = 0; c < MAXCH; ++c) { o4[c].x *= f; o4[c].y *= f; o4[c].z *= f; o4[c].w *= f; }
#pragma unroll
            for (int r = 0; r < RU2; ++r) {
                const float p = expf(e[r] - m_new);
                s_run += p;
#pragma unroll
                for (int c = 0; c < MAXCH; ++c) {
                    o4[c].x = fmaf(p, vv[r][c].x, o4[c].x); o4[c].y = fmaf(p, vv[r][c].y, o4[c].y);
                    o4[c].z = fmaf(p, vv[r][c].z, o4[c].z); o4[c].w = fmaf(p, vv[r][c].w, o4[c].w);
                }
            }
            m_run = m_new;
        } else {
            float* derow = a.de + ((long long)b * heads + h) * a.ld_w;
#pragma unroll
            for (int r = 0; r < RU2; ++r) {
                const int t = tb + r * NW2;
                const float de = wt[r] * (e[r] - dot) * a.scale;      // 0 for rows past t_hi (wt = 0)
                if (lane == 0 && t < t_hi) derow[t] = de;
#pragma unroll
                for (int c = 0; c < MAXCH; ++c) {
                    o4[c].x = fmaf(de, kk[r][c].x, o4[c].x); o4[c].y = fmaf(de, kk[r][c].y, o4[c].y);
                    o4[c].z = fmaf(de, kk[r][c].z, o4[c].z); o4[c].w = fmaf(de, kk[r][c].w, o4[c].w);
                }
            }
        }
    }
    // ---- fused tail: operands of the cell backward and the classifier-path dq, fetched behind the merges below ----
    float t_g[4] = {0.f, 0.f, 0.f, 0.f}, t_c = 0.f, t_cp = 0.f, t_dc = 0.f, t_mask = 1.f, t_rec = 0.f, t_dq_old = 0.f;
    const int tu = rank * nper + tid;                        // hidden unit of thread tid < nper
    if constexpr (TAIL) {
        if (tid < d && a.dq_accumulate) t_dq_old = a.dq[(long long)b * a.ld_dq + tid];     // read before any CTA of the pair writes dq
        if (tid < nper) {
            const int H = tl.DO;
            float pb[TAIL_MAX_SPLIT];
            const int nb = tl.dh_b ? (tl.nsplit_b > 1 ? tl.nsplit_b : 1) : 0;
#pragma unroll
            for (int sp = 0; sp < TAIL_MAX_SPLIT; ++sp) pb[sp] = sp < nb ? tl.dh_b[sp * tl.stride_b + (long long)b * tl.ld_b + tu] : 0.f;
            const float* g = tl.G + (long long)b * 4 * H + tu;
            t_g[0] = g[0]; t_g[1] = g[H]; t_g[2] = g[2 * H]; t_g[3] = g[3 * H];
            t_c = tl.c[(long long)b * tl.ld_c + tu]; t_cp = tl.c_prev[(long long)b * tl.ld_cp + tu];
            t_dc = tl.first ? 0.f : tl.dc[(long long)b * H + tu];
            if (tl.mask) t_mask = tl.mask[(long long)b * H + tu];
#pragma unroll
            for (int sp = 0; sp < TAIL_MAX_SPLIT; ++sp) t_rec += pb[sp];
        }
    }
    // ---- merge the warps of this CTA ----
    if (lane == 0) { wm[w] = m_run; wsum[w] = s_run; }
#pragma unroll
    for (int c = 0; c < MAXCH; ++c) {
        const int k = c * 128 + lane * 4;
        if (c < nch && k < d) *reinterpret_cast<float4*>(part + w * d + k) = o4[c];
    }
    __syncthreads();
    float m_c = -INFINITY;
    if (!BWD) {
#pragma unroll
        for (int i = 0; i < NW2; ++i) m_c = fmaxf(m_c, wm[i]);
    }
    for (int p = tid; p < d; p += NT2) {
        float r = 0.f;
#pragma unroll
        for (int i = 0; i < NW2; ++i) {
            const float f = BWD ? 1.f : ((wm[i] == -INFINITY) ? 0.f : expf(wm[i] - m_c));
            r = fmaf(f, part[i * d + p], r);
        }
        cta_o[p] = r;
    }
    if (!BWD && tid == 0) {
        float s_c = 0.f;
#pragma unroll
        for (int i = 0; i < NW2; ++i) s_c += (wm[i] == -INFINITY) ? 0.f : wsum[i] * expf(wm[i] - m_c);
        cta_ms[0] = m_c; cta_ms[1] = s_c;
    }
    cluster_sync_all();          // every CTA's (m, s, o) is in its shared memory and visible cluster-wide
    float M = 0.f, inv = 1.f;
    if (!BWD) {
        if (tid < 2 * S) cl[tid] = ld_dsmem_f32(cta_ms + (tid & 1), (unsigned)(tid >> 1));
        __syncthreads();
        M = -INFINITY;
        for (int j = 0; j < S; ++j) M = fmaxf(M, cl[2 * j]);
        float tot = 0.f;
        for (int j = 0; j < S; ++j) tot += (cl[2 * j] == -INFINITY) ? 0.f : cl[2 * j + 1] * expf(cl[2 * j] - M);
        inv = 1.f / tot;
        // normalised weights of this CTA's rows; exact zeros past the length (src/models.py:171-175)
        float* wrow = a.w + ((long long)b * heads + h) * a.ld_w;
        for (int t = t_lo + tid; t < t_hi; t += NT2) {
            const float wv = expf(sc[t - t_lo] - M) * inv;
            wrow[t] = wv;
            if (b == 0 && a.w_b0) a.w_b0[(long long)h * T + t] = wv;
        }
        for (int t = len + rank * NT2 + tid; t < T; t += S * NT2) {
            wrow[t] = 0.f;
            if (b == 0 && a.w_b0) a.w_b0[(long long)h * T + t] = 0.f;
        }
    } else {
        float* derow = a.de + ((long long)b * heads + h) * a.ld_w;
        for (int t = len + rank * NT2 + tid; t < T; t += S * NT2) derow[t] = 0.f;
        if (a.dctx2 && rank == 0 && w == 0) {       // every CTA read dctx before the cluster barrier above
#pragma unroll
            for (int c = 0; c < MAXCH; ++c) {
                const int k = c * 128 + lane * 4;
                if (c < nch && k < d) *reinterpret_cast<float4*>(a.dctx + (long long)b * a.ld_dctx + h * d + k) = v4[c];
            }
        }
    }
    const int dper = (d + S - 1) / S;
    if constexpr (TAIL) {
        // every CTA of the pair sums ALL d columns (it needs the whole dq vector for its slice of dq . Wq); it stores its own columns
        if (tid < d) {
            float r = 0.f;
            for (int j = 0; j < S; ++j) r += ld_dsmem_f32(cta_o + tid, (unsigned)j);
            const float tot = t_dq_old + r;
            dqs[tid] = tot;
            if (tid / dper == rank) {
                a.dq[(long long)b * a.ld_dq + tid] = tot;
                if (a.dq_bf16) ((__nv_bfloat16*)a.dq_bf16)[(long long)b * a.ld_dq_bf16 + tid] = __float2bfloat16(tot);
            }
        }
        cluster_sync_all();      // peers are done reading this CTA's shared memory; dqs is complete (the barrier is also a CTA barrier)
        cp_async_wait<0>();
        __syncthreads();         // every thread's part of the weight slice has landed
        const int pairs = nper / 2, kg = NT2 / pairs, kper = P / kg;
        {
            const int j = tid % pairs, g = tid / pairs;
            float a0 = 0.f, a1 = 0.f;
            const unsigned* wrow = reinterpret_cast<const unsigned*>(wq_s) + j;
#pragma unroll 8
            for (int k = g * kper; k < (g + 1) * kper; ++k) {
                const unsigned w2 = wrow[k * pairs];
                const float x = dqs[k];
                a0 = fmaf(x, __uint_as_float(w2 << 16), a0);
                a1 = fmaf(x, __uint_as_float(w2 & 0xffff0000u), a1);
            }
            gpart[g * nper + 2 * j] = a0; gpart[g * nper + 2 * j + 1] = a1;
        }
        __syncthreads();
        if (tid < nper) {
            // LSTMCell-1 backward of hidden unit tu (decoder.cu::cell_bwd_kernel with dh_a = dq . Wq)
            const int H = tl.DO;
            float dh = 0.f;
            for (int g = 0; g < kg; ++g) dh += gpart[g * nper + tid];
            dh = (dh + t_rec) * t_mask;
            const float gi = t_g[0], gf = t_g[1], gg = t_g[2], go = t_g[3];
            const float tc = tanhf(t_c);
            const float dct = fmaf(dh * go, 1.f - tc * tc, t_dc);
            const float d0 = dct * gg * gi * (1.f - gi), d1 = dct * t_cp * gf * (1.f - gf);
            const float d2 = dct * gi * (1.f - gg * gg), d3 = dh * tc * go * (1.f - go);
            float* g = tl.G + (long long)b * 4 * H + tu;
            g[0] = d0; g[H] = d1; g[2 * H] = d2; g[3 * H] = d3;
            __nv_bfloat16* gb = reinterpret_cast<__nv_bfloat16*>(tl.Gb) + (long long)b * 4 * H + tu;
            gb[0] = __float2bfloat16(d0); gb[H] = __float2bfloat16(d1); gb[2 * H] = __float2bfloat16(d2); gb[3 * H] = __float2bfloat16(d3);
            tl.dc[(long long)b * H + tu] = dct * gf;
        }
    } else {
    // ---- output columns [rank*dper, (rank+1)*dper) of this (row, head): sum the cluster's partial vectors over DSMEM ----
    for (int i = tid; i < dper; i += NT2) {
        const int p = rank * dper + i;
        if (p >= d) break;
        float r = 0.f;
        for (int j = 0; j < S; ++j) {
            const float f = BWD ? 1.f : ((cl[2 * j] == -INFINITY) ? 0.f : expf(cl[2 * j] - M));
            r = fmaf(f, ld_dsmem_f32(cta_o + p, (unsigned)j), r);
        }
        if (!BWD) {
            r *= inv;
            a.ctx[(long long)b * a.ld_ctx + h * d + p] = r;
            if (a.ctx2) a.ctx2[(long long)b * a.ld_ctx2 + h * d + p] = r;
            if (a.ctx2_bf16) ((__nv_bfloat16*)a.ctx2_bf16)[(long long)b * a.ld_ctx2_bf16 + h * d + p] = __float2bfloat16(r);
        } else {
            float* dq = a.dq + (long long)b * a.ld_dq + h * d + p;
            const float tot = (a.dq_accumulate ? *dq : 0.f) + r;
            *dq = tot;
            if (a.dq_bf16) ((__nv_bfloat16*)a.dq_bf16)[(long long)b * a.ld_dq_bf16 + h * d + p] = __float2bfloat16(tot);
        }
    }
    cluster_sync_all();          // keep this CTA's shared memory alive until every peer has read it
    }
}

int check(const LasAttnStep* a, bool bwd) {
    LAS_CHECK_ARG(a != nullptr, "attn_step: null descriptor");
    LAS_CHECK_ARG(a->B >= 1 && a->T >= 1 && a->P >= 4 && a->heads >= 1, "attn_step: bad dims B=%d T=%d P=%d heads=%d", a->B,
                  a->T, a->P, a->heads);
    LAS_CHECK_ARG(a->P % a->heads == 0, "attn_step: proj_dim %d %% heads %d != 0", a->P, a->heads);
    int d = a->P / a->heads;
    LAS_CHECK_ARG(d % 4 == 0 && d <= 128 * MAXCH, "attn_step: head dim %d must be a multiple of 4 and <= %d", d, 128 * MAXCH);
    LAS_CHECK_ARG(a->K && a->V && a->lens && a->w, "attn_step: null K/V/lens/w");
    LAS_CHECK_ARG(!a->fmask || a->w2, "attn_step: the init-force prior (fmask) needs w2 (second-softmax weights)");
    LAS_CHECK_ARG(a->dctx2_nsplit <= 8, "attn_step: at most 8 dctx2 partials");
    if (!bwd) {
        LAS_CHECK_ARG(a->q && a->ctx, "attn_step_fwd: null q/ctx");
        LAS_CHECK_ARG(a->ld_q % 4 == 0, "attn_step_fwd: ld_q must be a multiple of 4");
    } else {
        LAS_CHECK_ARG(a->dctx && a->dq && a->de, "attn_step_bwd: null dctx/dq/de");
        LAS_CHECK_ARG(a->ld_dctx % 4 == 0 && (!a->dctx2 || a->ld_dctx2 % 4 == 0), "attn_step_bwd: ld_dctx must be a multiple of 4");
    }
    return LAS_OK;
}

size_t smem_bytes(const LasAttnStep* a) { return sizeof(float) * ((size_t)((a->T + 3) & ~3) + NW + (size_t)NW * (a->P / a->heads)); }

template <bool BWD, int RU, bool KV16>
int launch_one(const LasAttnStep* a, size_t smem, cudaStream_t st) {
    if (smem > 48 * 1024) LAS_CUDA(cudaFuncSetAttribute(attn_step_kernel<BWD, RU, KV16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LAS_CUDA(las_launch(attn_step_kernel<BWD, RU, KV16>, dim3(a->B * a->heads), dim3(NT), smem, st, *a));
    return LAS_OK;
}
size_t tail_smem_floats(int P, int DO, int S) {
    const int nper = DO / S, kg = NT2 / (nper / 2);
    return (size_t)P + (size_t)kg * nper + (size_t)P * nper / 2;       // dqs + gpart + bf16 weight slice
}

template <bool BWD, bool KV16, bool TAIL>
int launch_split(const LasAttnStep* a, int S, cudaStream_t st, const LasAttnCellTail* tail = nullptr) {
    const int d = a->P / a->heads;
    const int cap = ((a->T + S - 1) / S + 3) & ~3;
    const size_t smem = sizeof(float) * ((size_t)cap + 2 * NW2 + 16 + 4 + d + (size_t)NW2 * d + (TAIL ? tail_smem_floats(a->P, tail->DO, S) : 0));
    if (smem > 48 * 1024) LAS_CUDA(cudaFuncSetAttribute(attn_step_split_kernel<BWD, KV16, TAIL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(S, a->B * a->heads, 1);
    cfg.blockDim = dim3(NT2, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = S; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = las_pdl_active() ? 2 : 1;
    LasAttnCellTail tl{};
    if (TAIL) tl = *tail;
    LAS_CUDA(cudaLaunchKernelEx(&cfg, attn_step_split_kernel<BWD, KV16, TAIL>, *a, tl));
    return LAS_OK;
}

// CTAs per (batch row, head).  Measured on B200 (scripts/bench_attn.py, graph-replayed launches, profiles/attn_split_sweep_r1.txt):
// B=96,T=200: S=1 8.9 us, S=2 8.0 us, S=3 11.3 us, S=4 11.1 us, S=8 14.6 us (fwd; clusters larger than a CTA pair cost more
// to schedule than the extra SMs give back); B=256,T=375: S=1 31.2 us (6.3 TB/s), S=2 33.6 us.  So: a pair when the rows
// alone cannot fill the SMs and each half still has >= 32 rows, else one CTA per row.
int split_factor(const LasAttnStep* a) {
    const char* e = getenv("LAS_ATTN_SPLIT");       // tuning / test override (0 = two-phase kernel)
    const int forced = (e && *e) ? atoi(e) : -1;
    if (forced >= 0) return forced > 8 ? 8 : forced;
    const int rows = a->B * a->heads;
    return (rows <= las_device_info()->num_sms && a->T >= 64) ? 2 : 1;
}

template <bool BWD>
int launch_attn(const LasAttnStep* a, size_t smem, cudaStream_t st) {
    // single-pass T-split kernel; backward needs the saved context for it (dot = dctx . ctx).  LAS_ATTN_SPLIT=0 or a
    // backward descriptor without ctx selects the two-phase one-CTA-per-row kernel.
    const int S = ((BWD && !a->ctx) || a->fmask) ? 0 : split_factor(a);
    if (S >= 1) return a->kv_bf16 ? launch_split<BWD, true, false>(a, S, st) : launch_split<BWD, false, false>(a, S, st);
    const bool big = a->B * a->heads >= las_device_info()->num_sms;
    if (a->kv_bf16) return big ? launch_one<BWD, 4, true>(a, smem, st) : launch_one<BWD, 8, true>(a, smem, st);
    return big ? launch_one<BWD, 4, false>(a, smem, st) : launch_one<BWD, 8, false>(a, smem, st);
}

}  // namespace

extern "C" int las_attn_step_fwd_f32(const LasAttnStep* a, void* stream) {
    int rc = check(a, false);
    if (rc) return rc;
    rc = las_set_device_of(a->K);
    if (rc) return rc;
    size_t smem = smem_bytes(a);
    LasProfScope prof(LAS_PROF_ATTN_FWD, stream, 2.0 * a->B * (double)a->T * a->P * (a->kv_bf16 ? 2 : 4));
    // rows in flight per warp: 8 when the grid leaves SMs idle (train: 96 CTAs), 4 when it over-subscribes them (greedy: 256 CTAs)
    rc = launch_attn<false>(a, smem, (cudaStream_t)stream);
    if (rc) return rc;
    LAS_LAUNCH_CHECK();
    return LAS_OK;
}

extern "C" int las_attn_step_bwd_f32(const LasAttnStep* a, void* stream) {
    int rc = check(a, true);
    if (rc) return rc;
    rc = las_set_device_of(a->K);
    if (rc) return rc;
    size_t smem = smem_bytes(a);
    LasProfScope prof(LAS_PROF_ATTN_BWD, stream, 2.0 * a->B * (double)a->T * a->P * (a->kv_bf16 ? 2 : 4));
    rc = launch_attn<true>(a, smem, (cudaStream_t)stream);
    if (rc) return rc;
    LAS_LAUNCH_CHECK();
    return LAS_OK;
}

namespace {

// ---------------------------------------------------------------------------------------------------------------------
// Backward decoder step on the tensor pipe (fp16 K / V rows): one CTA of 8 warps per batch row, single head.
//   dw_t = dctx . V_t            -> mma.sync m16n8k16: A = 16 V rows x 16 columns, B = dctx (replicated over the 8 output columns)
//   de_t = w_t (dw_t - dctx.ctx) sqrt(d)
//   dq   = sum_t de_t K_t        -> mma.sync: A = K^T block (16 columns x 16 positions, paired up in registers with a byte permute),
//                                   B = de of the 16 positions (replicated)
// Same fragment tricks as the forward passes in decoder_persist.cu (16 contiguous bytes per lane and row; a dot product does not
// care how its terms are paired).  Gradients have no fixed range: dctx is scaled by a power of two per row so that its largest
// element is in [0.5, 1) before it is rounded to fp16 -- de then stays far from both ends of the fp16 range -- and dq / the stored de
// are scaled back in fp32.  A warp loads BOTH its V block and its K block before the first mma (one L2 round trip per 16 positions:
// 128 + 64 registers of operands and accumulators, hence one CTA per SM), then the fused tail of attn_tail.h runs on the whole
// 256-column dq in this CTA: no cluster, no DSMEM merge.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_f16_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
    const __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&h);
}

constexpr int BT_NT = 256, BT_NW = BT_NT / 32;

__global__ void __launch_bounds__(BT_NT, 1) attn_bwd_tc_kernel(LasAttnStep a, LasAttnCellTail tl) {
    extern __shared__ __align__(16) float sm[];
    const int T = a.T, P = a.P;                    // single head: d == P, P in {64, 128, 192, 256}
    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int g = lane >> 2, c = lane & 3;
    const int DO = tl.DO;
    float* part = sm;                              // [BT_NW][P]  per-warp partial dq
    float* dqs = part + BT_NW * P;                 // [P]         dq_total of the row
    float* red = dqs + P;                          // [2 * BT_NW] block reductions
    float* gpart = red + 2 * BT_NW;                // [kg][DO]    partial dh1 sums of the k groups
    const int pairs = DO / 2, kg = BT_NT / pairs, kper = P / kg;
    __half* dxh = reinterpret_cast<__half*>(gpart + kg * DO);     // [P] scaled dctx as fp16 (B operand of the first mma)
    __nv_bfloat16* wq_s = reinterpret_cast<__nv_bfloat16*>(dxh + ((P + 7) & ~7));      // [P][DO] query_map.weight, bf16
    {
        // the weights do not depend on the predecessor kernel: fetched before the programmatic-launch wait, asynchronously
        const int cpr = DO / 8;
        const __nv_bfloat16* wsrc = reinterpret_cast<const __nv_bfloat16*>(tl.wq_bf16);
        for (int i = tid; i < P * cpr; i += BT_NT) cp_async16(wq_s + i * 8, wsrc + (long long)i * 8);
        cp_async_commit();
    }
    pdl_wait();
    pdl_trigger();
    const int len = min(a.lens[b], T);

    // ---- dctx_total (classifier path + cell-0 path of the next step, split-K partials), its scale, dctx . ctx ----
    float dct = 0.f, dq_old = 0.f;
    if (tid < P) {
        dct = a.dctx[(long long)b * a.ld_dctx + tid];
        if (a.dctx2) {
            float e[8];
            const int ns = a.dctx2_nsplit > 1 ? a.dctx2_nsplit : 1;
#pragma unroll
            for (int sp = 0; sp < 8; ++sp) e[sp] = sp < ns ? a.dctx2[sp * a.dctx2_split_stride + (long long)b * a.ld_dctx2 + tid] : 0.f;
#pragma unroll
            for (int sp = 0; sp < 8; ++sp) dct += e[sp];
            a.dctx[(long long)b * a.ld_dctx + tid] = dct;          // the batched dV GEMM after the loop reads the total
        }
        if (a.dq_accumulate) dq_old = a.dq[(long long)b * a.ld_dq + tid];
    }
    float amax = warp_max(fabsf(dct));
    float dot = warp_sum(tid < P ? dct * a.ctx[(long long)b * a.ld_ctx + tid] : 0.f);
    if (lane == 0) { red[w] = amax; red[BT_NW + w] = dot; }
    // ---- operands of the cell backward (thread = hidden unit), in flight behind everything below ----
    float t_g[4] = {0.f, 0.f, 0.f, 0.f}, t_c = 0.f, t_cp = 0.f, t_dc = 0.f, t_mask = 1.f, t_rec = 0.f;
    if (tid < DO) {
        float pb[TAIL_MAX_SPLIT];
        const int nb = tl.dh_b ? (tl.nsplit_b > 1 ? tl.nsplit_b : 1) : 0;
#pragma unroll
        for (int sp = 0; sp < TAIL_MAX_SPLIT; ++sp) pb[sp] = sp < nb ? tl.dh_b[sp * tl.stride_b + (long long)b * tl.ld_b + tid] : 0.f;
        const float* gp = tl.G + (long long)b * 4 * DO + tid;
        t_g[0] = gp[0]; t_g[1] = gp[DO]; t_g[2] = gp[2 * DO]; t_g[3] = gp[3 * DO];
        t_c = tl.c[(long long)b * tl.ld_c + tid]; t_cp = tl.c_prev[(long long)b * tl.ld_cp + tid];
        t_dc = tl.first ? 0.f : tl.dc[(long long)b * DO + tid];
        if (tl.mask) t_mask = tl.mask[(long long)b * DO + tid];
#pragma unroll
        for (int sp = 0; sp < TAIL_MAX_SPLIT; ++sp) t_rec += pb[sp];
    }
    __syncthreads();
    amax = red[0]; dot = red[BT_NW];
#pragma unroll
    for (int i = 1; i < BT_NW; ++i) { amax = fmaxf(amax, red[i]); dot += red[BT_NW + i]; }
    int ex = 0;
    if (amax > 0.f && amax < 3.0e38f) frexpf(amax, &ex);
    const float sc_up = ldexpf(1.f, -ex), sc_dn = ldexpf(1.f, ex);       // amax * sc_up in [0.5, 1)
    if (tid < P) dxh[tid] = __float2half_rn(dct * sc_up);
    const float dot_s = dot * sc_up;
    __syncthreads();

    const __half* Kh = reinterpret_cast<const __half*>(tl.K_f16) + (long long)b * T * P;
    const __half* Vh = reinterpret_cast<const __half*>(tl.V_f16) + (long long)b * T * P;
    const float* wrow_c = a.w + (long long)b * a.ld_w;
    float* derow = a.de + (long long)b * a.ld_w;
    const int nchunk = P / 64 > 0 ? (P + 63) / 64 : 1;
    float acc[4][4][4];
#pragma unroll
    for (int ch = 0; ch < 4; ++ch)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[ch][j][0] = 0.f; acc[ch][j][1] = 0.f; acc[ch][j][2] = 0.f; acc[ch][j][3] = 0.f; }

    for (int t0 = w * 16; t0 < len; t0 += BT_NW * 16) {
        // V block, "row" pattern: lane (g, c) holds rows t0+g / t0+g+8, columns 32 pb + 8c .. + 8
        const int ta = t0 + g, tb = t0 + g + 8;
        const bool va = ta < len, vb = tb < len;
        uint4 xa[8], xb[8];
#pragma unroll
        for (int pb = 0; pb < 8; ++pb) {
            const bool in = pb * 32 < P;
            xa[pb] = (in && va) ? __ldg(reinterpret_cast<const uint4*>(Vh + (long long)ta * P + pb * 32 + c * 8)) : make_uint4(0u, 0u, 0u, 0u);
            xb[pb] = (in && vb) ? __ldg(reinterpret_cast<const uint4*>(Vh + (long long)tb * P + pb * 32 + c * 8)) : make_uint4(0u, 0u, 0u, 0u);
        }
        // K block, "transposed" pattern: lane (g, c) holds positions t0+2c, +1, t0+8+2c, +1, columns 64 ch + 8g .. + 8
        const int tt[4] = {t0 + 2 * c, t0 + 2 * c + 1, t0 + 8 + 2 * c, t0 + 9 + 2 * c};
        uint4 r[4][4];
#pragma unroll
        for (int ch = 0; ch < 4; ++ch)
#pragma unroll
            for (int i = 0; i < 4; ++i)
                r[ch][i] = (ch < nchunk && tt[i] < len) ? __ldg(reinterpret_cast<const uint4*>(Kh + (long long)tt[i] * P + ch * 64 + 8 * g))
                                                        : make_uint4(0u, 0u, 0u, 0u);
        const float wa_ = va ? wrow_c[ta] : 0.f, wb_ = vb ? wrow_c[tb] : 0.f;
        float d[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int pb = 0; pb < 8; ++pb) {
            if (pb * 32 < P) {
                const uint4 qq = *reinterpret_cast<const uint4*>(dxh + pb * 32 + c * 8);
                mma_f16_16816(d, xa[pb].x, xb[pb].x, xa[pb].y, xb[pb].y, qq.x, qq.y);
                mma_f16_16816(d, xa[pb].z, xb[pb].z, xa[pb].w, xb[pb].w, qq.z, qq.w);
            }
        }
        // every lane of row group g holds dw (scaled) of rows t0+g (d[0]) and t0+g+8 (d[2])
        const float de_a = wa_ * (d[0] - dot_s) * a.scale, de_b = wb_ * (d[2] - dot_s) * a.scale;      // 0 past the length (w = 0)
        if (c == 0) {
            if (va) derow[ta] = de_a * sc_dn;
            if (vb) derow[tb] = de_b * sc_dn;
        }
        // B fragment of the second mma: positions 2c, 2c+1 (rows held by lane groups g = 2c, 2c+1) and 8+2c, 9+2c
        const uint32_t fb0 = pack_h2(__shfl_sync(0xffffffffu, de_a, (2 * c) * 4), __shfl_sync(0xffffffffu, de_a, (2 * c + 1) * 4));
        const uint32_t fb1 = pack_h2(__shfl_sync(0xffffffffu, de_b, (2 * c) * 4), __shfl_sync(0xffffffffu, de_b, (2 * c + 1) * 4));
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
            if (ch < nchunk) {
                const uint32_t r0[4] = {r[ch][0].x, r[ch][0].y, r[ch][0].z, r[ch][0].w}, r1[4] = {r[ch][1].x, r[ch][1].y, r[ch][1].z, r[ch][1].w};
                const uint32_t r2[4] = {r[ch][2].x, r[ch][2].y, r[ch][2].z, r[ch][2].w}, r3[4] = {r[ch][3].x, r[ch][3].y, r[ch][3].z, r[ch][3].w};
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    mma_f16_16816(acc[ch][j], __byte_perm(r0[j], r1[j], 0x5410u), __byte_perm(r0[j], r1[j], 0x7632u),
                                  __byte_perm(r2[j], r3[j], 0x5410u), __byte_perm(r2[j], r3[j], 0x7632u), fb0, fb1);
            }
        }
    }
    // output columns of D are identical: lanes with c == 0 hold columns 64 ch + 8g + 2j (acc[..][0]) and + 2j + 1 (acc[..][2])
    if (c == 0) {
#pragma unroll
        for (int ch = 0; ch < 4; ++ch)
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (ch * 64 + 8 * g + 2 * j < P) {
                    part[w * P + ch * 64 + 8 * g + 2 * j] = acc[ch][j][0];
                    part[w * P + ch * 64 + 8 * g + 2 * j + 1] = acc[ch][j][2];
                }
    }
    for (int t = len + tid; t < T; t += BT_NT) derow[t] = 0.f;
    __syncthreads();
    if (tid < P) {
        float rsum = 0.f;
#pragma unroll
        for (int i = 0; i < BT_NW; ++i) rsum += part[i * P + tid];
        const float tot = dq_old + rsum * sc_dn;
        dqs[tid] = tot;
        a.dq[(long long)b * a.ld_dq + tid] = tot;
        if (a.dq_bf16) ((__nv_bfloat16*)a.dq_bf16)[(long long)b * a.ld_dq_bf16 + tid] = __float2bfloat16(tot);
    }
    cp_async_wait<0>();
    __syncthreads();
    // ---- dh1 = dq_total . Wq from shared memory, then LSTMCell-1 backward (same arithmetic as the TAIL branch of attn_step_split_kernel) ----
    {
        const int j = tid % pairs, gk = tid / pairs;
        float a0 = 0.f, a1 = 0.f;
        const unsigned* wr = reinterpret_cast<const unsigned*>(wq_s) + j;
#pragma unroll 8
        for (int k = gk * kper; k < (gk + 1) * kper; ++k) {
            const unsigned w2 = wr[k * pairs];
            const float x = dqs[k];
            a0 = fmaf(x, __uint_as_float(w2 << 16), a0);
            a1 = fmaf(x, __uint_as_float(w2 & 0xffff0000u), a1);
        }
        gpart[gk * DO + 2 * j] = a0; gpart[gk * DO + 2 * j + 1] = a1;
    }
    __syncthreads();
    if (tid < DO) {
        float dh = 0.f;
        for (int gk = 0; gk < kg; ++gk) dh += gpart[gk * DO + tid];
        dh = (dh + t_rec) * t_mask;
        const float gi = t_g[0], gf = t_g[1], gg = t_g[2], go = t_g[3];
        const float tc = tanhf(t_c);
        const float dctt = fmaf(dh * go, 1.f - tc * tc, t_dc);
        const float d0 = dctt * gg * gi * (1.f - gi), d1 = dctt * t_cp * gf * (1.f - gf);
        const float d2 = dctt * gi * (1.f - gg * gg), d3 = dh * tc * go * (1.f - go);
        float* gp = tl.G + (long long)b * 4 * DO + tid;
        gp[0] = d0; gp[DO] = d1; gp[2 * DO] = d2; gp[3 * DO] = d3;
        __nv_bfloat16* gb = reinterpret_cast<__nv_bfloat16*>(tl.Gb) + (long long)b * 4 * DO + tid;
        gb[0] = __float2bfloat16(d0); gb[DO] = __float2bfloat16(d1); gb[2 * DO] = __float2bfloat16(d2); gb[3 * DO] = __float2bfloat16(d3);
        tl.dc[(long long)b * DO + tid] = dctt * gf;
    }
}

}  // namespace

// ---- fused backward step (attn_tail.h) ----
static int tail_split(const LasAttnStep* a, int DO) {
    if (!a || a->heads != 1 || a->fmask || !a->ctx || a->kv_bf16 || a->P > NT2 || DO < 16) return 0;
    const int S = split_factor(a);
    if (S < 1 || S > 2 || DO % (8 * S) != 0) return 0;
    const int nper = DO / S, pairs = nper / 2;
    if (pairs > NT2 || NT2 % pairs != 0 || a->P % (NT2 / pairs) != 0 || nper > NT2) return 0;
    const int cap = ((a->T + S - 1) / S + 3) & ~3;
    const size_t smem = sizeof(float) * ((size_t)cap + 2 * NW2 + 16 + 4 + a->P + (size_t)NW2 * a->P + tail_smem_floats(a->P, DO, S));
    return smem <= (S == 2 ? 100 : 200) * 1024 ? S : 0;          // pairs: two CTAs per SM stay resident; one CTA per row: at most one per SM anyway
}

int las_attn_step_bwd_cell_supported(const LasAttnStep* a, int DO) {
    const char* e = getenv("LAS_BWD_FUSE_TAIL");
    if (e && *e == '0') return 0;
    return tail_split(a, DO) > 0 ? 1 : 0;
}

static size_t bwd_tc_smem(int P, int DO) {
    const int kg = BT_NT / (DO / 2);
    return sizeof(float) * ((size_t)BT_NW * P + P + 2 * BT_NW + (size_t)kg * DO) + 2 * (size_t)((P + 7) & ~7) + 2 * (size_t)P * DO + 16;
}
// tensor-pipe form: fp16 K / V copies given, single head, one thread per hidden unit and per attention column
static bool bwd_tc_ok(const LasAttnStep* a, const LasAttnCellTail* tail) {
    const char* e = getenv("LAS_BWD_ATT_TC");
    if (e && *e == '0') return false;
    if (!tail->K_f16 || !tail->V_f16 || a->heads != 1 || a->fmask || !a->ctx || a->kv_bf16) return false;
    const int P = a->P, DO = tail->DO;
    if (P % 64 != 0 || P > 256 || DO % 16 != 0 || DO > BT_NT || DO < 32) return false;
    const int pairs = DO / 2;
    if (BT_NT % pairs != 0 || P % (BT_NT / pairs) != 0) return false;
    return bwd_tc_smem(P, DO) <= 200 * 1024;
}

int las_attn_step_bwd_cell(const LasAttnStep* a, const LasAttnCellTail* tail, void* stream) {
    int rc = check(a, true);
    if (rc) return rc;
    if (tail && bwd_tc_ok(a, tail)) {
        LAS_CHECK_ARG(tail->wq_bf16 && tail->G && tail->Gb && tail->c && tail->c_prev && tail->dc, "attn_step_bwd_cell: null tail operand");
        LAS_CHECK_ARG(tail->nsplit_b <= TAIL_MAX_SPLIT, "attn_step_bwd_cell: at most %d recurrent partials", TAIL_MAX_SPLIT);
        rc = las_set_device_of(a->K);
        if (rc) return rc;
        LasProfScope prof(LAS_PROF_ATTN_BWD, stream, 2.0 * a->B * (double)a->T * a->P * 2);
        const size_t smem = bwd_tc_smem(a->P, tail->DO);
        LAS_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(a->B, 1, 1);
        cfg.blockDim = dim3(BT_NT, 1, 1);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = (cudaStream_t)stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = las_pdl_active() ? 1 : 0;
        LAS_CUDA(cudaLaunchKernelEx(&cfg, attn_bwd_tc_kernel, *a, *tail));
        LAS_LAUNCH_CHECK();
        return LAS_OK;
    }
    LAS_CHECK_ARG(tail && tail->wq_bf16 && tail->G && tail->Gb && tail->c && tail->c_prev && tail->dc, "attn_step_bwd_cell: null tail operand");
    LAS_CHECK_ARG(tail->nsplit_b <= TAIL_MAX_SPLIT, "attn_step_bwd_cell: at most %d recurrent partials", TAIL_MAX_SPLIT);
    const int S = tail_split(a, tail->DO);
    LAS_CHECK_ARG(S > 0, "attn_step_bwd_cell: shape not supported (ask las_attn_step_bwd_cell_supported)");
    rc = las_set_device_of(a->K);
    if (rc) return rc;
    LasProfScope prof(LAS_PROF_ATTN_BWD, stream, 2.0 * a->B * (double)a->T * a->P * 4);
    rc = launch_split<true, false, true>(a, S, (cudaStream_t)stream, tail);
    if (rc) return rc;
    LAS_LAUNCH_CHECK();
    return LAS_OK;
}
