// Internal (not part of the C ABI): the backward attention step of the Speller loop with its two dependents fused behind it.
//
// One backward decoder step of the reference (autograd of src/models.py:352-380) is the chain
//     attention backward (dq) -> dh1 = dq_total . Wq -> LSTMCell-1 backward (pointwise) -> dG1 . Wcat1 -> LSTMCell-0 backward -> dG0 . Wcat0
// The first three links are local to ONE batch row: the CTA pair that owns the row's attention already holds dq_total, a
// (P x DO) bf16 query_map weight slice fits its shared memory (fetched with cp.async while K / V stream), and the cell-1 pointwise
// backward needs only that row's saved gates.  Fusing them removes two dependent launches (a 96 x 256 x 256 GEMM and a pointwise
// kernel) from every step of the loop.
#pragma once
#include "las_b200.h"

struct LasAttnCellTail {
    const void* wq_bf16;      // (P, DO) row-major bf16: query_map.weight (dh1[n] = sum_k dq[k] Wq[k][n])
    int DO;
    // LSTMCell-1 backward of the same step (same meaning as decoder.cu's CellBwd; dh_a is the fused dq . Wq)
    float* G;                 // (B, 4*DO) in: activated gates, out: d(pre-activation)
    const float* dh_b;        // nullable: recurrent path, nsplit_b un-reduced split-K partials stride_b floats apart
    long long ld_b, stride_b;
    int nsplit_b;
    const float* mask;        // nullable (B, DO) dropout mask of h1
    const float* c; long long ld_c;
    const float* c_prev; long long ld_cp;
    float* dc;                // (B, DO) carried in / out
    void* Gb;                 // (B, 4*DO) bf16 copy of d(pre-activation)
    int first;                // 1: carried-in dc is zero
    // optional IEEE fp16 copies of K / V (B, T, P): the attention part then runs on the tensor pipe (attn_bwd_tc_kernel)
    const void* K_f16;
    const void* V_f16;
};

// 1 when las_attn_step_bwd_cell can run this step fused (single head, T-split kernel, weight slice fits shared memory)
int las_attn_step_bwd_cell_supported(const LasAttnStep* a, int DO);
int las_attn_step_bwd_cell(const LasAttnStep* a, const LasAttnCellTail* tail, void* stream);
