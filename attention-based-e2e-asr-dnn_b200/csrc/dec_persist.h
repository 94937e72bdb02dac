// Internal interface between decoder.cu (host-side loop driver, workspace layout) and decoder_persist.cu (the persistent
// decoder-step kernels).  Not part of the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>

// Forward loop of Speller.forward (reference src/models.py:336-385 with src/modules.py:340-365 inlined) as ONE kernel launch.
// All history arrays have `hist` (training: steps + 1, eval: 2) or `ghist` (training: steps, eval: 1) slots.
struct LasDecPersistFwd {
    int B, T, P, DH, DO, V, steps, training, sos_idx;
    int nsl;                       // batch slices of 32 rows
    int ngroups;                   // slice groups: cell CTA (r, sg) handles slices sg, sg + ngroups, ... one after the other
    int hist, ghist;
    int per_step_logits;           // 1: classifier + argmax inside the loop (greedy decoding, teacher forcing rate < 1)
    int kv16;                      // 1: K / V hold fp16
    float scale;                   // sqrt(P / heads)   (reference src/models.py:93,170: e = q.K / norm_factor)
    // parameters
    const float* emb;              // (V, 2P) tied classifier weight
    const float* cls_b;            // (V)
    const float* bq;               // (P)
    const float* b_ih1; const float* b_hh1;   // (4DO)
    const float* init_query;       // (DO)
    const float* Gemb;             // (V, 4DH) emb . W_ih0[:, :E]^T + b_ih0 + b_hh0
    const __half* W0;              // (4DH, P + DH) fp16 [W_ih0[:, E:] | W_hh0]
    const __half* W1;              // (4DO, DH + DO) fp16 [W_ih1 | W_hh1]
    const __half* WqT;             // (DO, P) fp16, query_map.weight transposed
    // inputs
    const void* K; const void* Vv; // (B, T, P) fp32 or fp16
    const int* enc_lens;           // (B)
    const int* y; long long ld_y;  // gold tokens (B, >= steps) or null
    const int* use_gold;           // (steps) device flags or null: step t > 0 feeds y[:, t-1] when use_gold[t] != 0
    const float* drop0; const float* drop1;   // (steps, B, DH) / (steps, B, DO) dropout masks or null
    // history / outputs
    __half* S0h;                   // (hist, B, P + DH) fp16 cell-0 operand rows [ctx_t | h0_{t-1}]
    __half* S1h;                   // (hist, B, DH + DO) fp16 cell-1 operand rows [h0_t | h1_{t-1}]
    __nv_bfloat16* S0b; __nv_bfloat16* S1b;   // bf16 copies of the same rows for backward's weight-gradient GEMMs (training) or null
    float* C0; float* C1;          // (hist, B, DH) / (hist, B, DO)
    float* G0; float* G1;          // (ghist, B, 4DH) / (ghist, B, 4DO) activated gates (training)
    float* QC;                     // (hist, B, 2P) [q | ctx]
    float* W;                      // (hist, B, T) attention weights
    float* att0;                   // (steps + 1, T) attention weights of batch row 0, or null
    float* logits;                 // (B, steps, V) (written here only when per_step_logits)
    int* chars;                    // (steps, B) argmax (per_step_logits)
    int* tok;                      // (steps, B) token fed at each step (training) or null
    unsigned* ctr;                 // 3 * nsl counters, 32 words apart, zeroed before the launch
    unsigned* err;                 // 1 word, zeroed: set when a hand-off wait times out (the kernel then traps)
    long long* dbg;                // optional (256 steps x 16 slots) globaltimer stamps, see las_dec_persist_set_debug
};

int las_dec_persist_fwd_supported(int B, int T, int P, int DH, int DO, int V, int heads, int init_force);
size_t las_dec_persist_ctr_words(int B);
int las_dec_persist_fwd_launch(const LasDecPersistFwd* a, cudaStream_t st);
