/*
 * las_b200 -- C ABI of the B200-native Listen-Attend-Spell hot path.
 *
 * The reference (Astromsoc/attention-based-e2e-asr-dnn) is pure Python/PyTorch and has NO native / FFI interface; the
 * drop-in boundary a user sees is the nn.Module API of src/models.py / src/modules.py (kept by the Python package
 * las_b200).  This header is the boundary underneath it: the entry points our modules bind with
 * ctypes, each citing the reference call site (file:line under /root/reference) whose PyTorch library dispatch it
 * replaces.  Conventions:
 *   - extern "C", plain pointers and sizes; every pointer is a DEVICE pointer unless the name ends in _host;
 *   - the caller (PyTorch) owns every buffer, including workspaces (size queries: *_workspace_*);
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises the host;
 *   - return 0 on success, negative LAS_ERR_* otherwise; las_last_error() gives the message (thread-local);
 *   - no CPU fallback: an entry either launches sm_100a kernels or returns an error;
 *   - re-entrant across devices (one process per GPU); the device of the first pointer argument is made current,
 *     so calls from the autograd engine thread are safe.
 */
#ifndef LAS_B200_H
#define LAS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LAS_B200_ABI_VERSION 1

/* ---- runtime ------------------------------------------------------------------------------------------------ */
int las_abi_version(void);
/* idempotent per-device initialisation (SM count, shared-memory opt-in, cooperative-launch check) */
int las_init(int device);
const char* las_last_error(void);
/* number of kernels this library has launched on the calling thread's devices since the last reset
 * (bench.py's gpu_launches) */
long long las_launch_count(void);
void las_launch_count_reset(void);

/* per-kernel-kind device timing for bench.py: CUDA events recorded around launches of the enabled kinds on the
 * launching stream.  kinds: 0 gate GEMMs, 1 other GEMMs, 2 recurrence fwd, 3 recurrence bwd, 4 attention step fwd,
 * 5 attention step bwd, 6 optimizer, 7 speller loop fwd, 8 speller loop bwd.  Synchronise before collecting.
 * `total_work` = algorithmic FLOPs (GEMM kinds), bytes (attention kinds) or timesteps (recurrence / speller kinds). */
void las_prof_enable(unsigned kind_mask);
void las_prof_reset(void);
int las_prof_collect(int kind, double* total_ms, long long* count, double* total_work);

/* ---- fp32 GEMM with two-level strided indexing ------------------------------------------------------------------
 * C[m][n] = alpha * sum_k A(m,k) B(k,n) + beta*C[m][n] + bias1[n] + bias2[n]
 * index i -> (i / inner) * s_outer + (i % inner) * s_inner ; inner == 0 means single level (i * s_inner).
 * Replaces: nn.LSTM's input projection X.W_ih^T (src/modules.py:80,189), the pyramidal reshape feeding it
 * (src/modules.py:171-185, done here as addressing), nn.Linear key/value/query maps (src/models.py:143-149,166),
 * the LSTMCell GEMMs (src/modules.py:355), the tied classifier (src/models.py:373), and every autograd matmul of
 * those in backward. */
typedef struct {
    const float* A;
    const float* B;
    float* C;
    const float* bias1; /* nullable, length N */
    const float* bias2; /* nullable, length N */
    int M, N, K, batch;
    long long a_m_so, a_m_si; int a_m_inner;
    long long a_k_so, a_k_si; int a_k_inner;
    long long b_k_so, b_k_si; int b_k_inner;
    long long b_n_s;
    long long c_m_so, c_m_si; int c_m_inner; /* C's n stride is 1 */
    long long bsA, bsB, bsC;                 /* batch strides (elements) */
    float alpha, beta;
    int prof_tag;                            /* 1: count this launch as an LSTM input-gate GEMM in las_prof_*; 2: ... one that runs beside a
                                              * recurrence kernel on the SMs it leaves free (timed apart: its duration is not a whole-GPU figure) */
} LasGemmF32;
int las_gemm_f32(const LasGemmF32* desc, void* stream);

/* ---- bf16 tensor-core GEMM (tcgen05.mma + TMEM + TMA), fp32 accumulate / fp32 output ---------------------------------
 * The tensor-pipe path of the LSTM input-gate projections (same reference call sites as las_gemm_f32's gate uses:
 * src/modules.py:80,189 forward; autograd dX / dW_ih / dW_hh in backward).  Operands are bf16; each is a 3-D tensor
 * (contiguous dim, rows with stride s1, batches with stride s2) handed to TMA, so the pyramid reshape-concat, odd-frame
 * drop and padding of src/modules.py:171-185 are tensor-map geometry, not copies.
 *   a_mn_major = 0 (forward / dgrad):  A[b][r][k] = A + b*a_s2 + r*a_s1 + k , r < M rows per batch, a_batches batches
 *        b_mn_major = 0: B[n][k] = B + n*b_s1 + k      (C = A . B^T, weights as stored)
 *        b_mn_major = 1: B[k][n] = B + k*b_s1 + n      (C = A . B,   weights as stored, dgrad)
 *        C row (b, r) at C + b*c_bs + r*ldc ; `lens` (a_batches, nullable) skips 128-row tiles starting at r >= lens[b]
 *   a_mn_major = 1 (weight gradient):  A[kb][k][m] = A + kb*a_s2 + k*a_s1 + m, B[kb][k][n] = B + kb*b_s2 + k*b_s1 + n,
 *        K rows per K-batch, k_batches batches;  C[m][n] = sum_{kb,k} A*B at C + m*ldc + n
 * bias1/bias2 (length N, nullable) are added in the epilogue; accumulate != 0 adds into C.  Split-K partials are reduced by
 * a second, deterministic kernel. */
typedef struct {
    const void* A; const void* B; float* C;
    const float* bias1; const float* bias2;
    int M, N, K;
    int a_batches, k_batches;
    long long a_s1, a_s2, b_s1, b_s2;
    long long c_bs, ldc;
    int a_mn_major, b_mn_major, accumulate;
    const int* lens;
    int prof_tag;
    double prof_flops;     /* algorithmic FLOPs credited to this launch by las_prof_* (0: 2*M*N*K*batches as launched) */
    int splitk;            /* weight-gradient form only: > 1 splits the reduction over that many CTAs per output tile */
    float* workspace;      /* split-K partial sums, >= splitk * M * round_up(N, 4) floats */
    int max_ctas;          /* > 0: launch at most this many (persistent) CTAs -- for a GEMM that runs on a second stream beside a
                            * recurrence kernel and should take only the SMs that kernel leaves free; 0: one CTA per SM */
    int a_f16, b_f16;      /* != 0: the operands hold IEEE fp16 instead of bf16 (bounded-range forward activations: 3 more mantissa
                            * bits); both flags must agree -- a mixed fp16 x bf16 pair is an illegal instruction on B200 */
    int k_chunk;           /* > 0 (a_mn_major = b_mn_major = 0 only): the reduction runs over K / k_chunk chunks of k_chunk elements whose */
    long long k_chunk_stride; /* starts lie k_chunk_stride elements apart in BOTH operands' K dimension.  One direction's half of a BiLSTM
                            * layer's output (src/modules.py:80: [h_fwd | h_bwd] per frame, two frames per row under the pyramid concat of
                            * :171-185) times the matching columns of the next layer's W_ih: that half of the gate projection can start as
                            * soon as ITS sweep has passed a frame, without waiting for the other direction */
} LasGemmTc;
int las_gemm_bf16_tc(const LasGemmTc* desc, void* stream);
/* dst[r][c] (bf16, row stride ld_dst) = c < cols ? srcrow(r)[c] : 0, c < cols_pad; source row r starts at
 * (r / inner)*bs + (r % inner)*ld_src when inner > 0 (a (B,T,F) view with a batch stride), else at r*ld_src */
int las_cast_f32_to_bf16(const float* src, long long ld_src, long long inner, long long bs, void* dst, long long ld_dst,
                         long long rows, int cols, int cols_pad, void* stream);
/* same with an IEEE fp16 destination */
int las_cast_f32_to_f16(const float* src, long long ld_src, long long inner, long long bs, void* dst, long long ld_dst,
                        long long rows, int cols, int cols_pad, void* stream);

/* column sums: out[n] (+)= sum_m X[m*ld + n], m < M, n < N.  Bias gradients (autograd of the bias adds in nn.LSTM /
 * nn.LSTMCell / nn.Linear). `scratch` needs las_colsum_scratch_floats(N) floats. */
size_t las_colsum_scratch_floats(int N);
int las_colsum_f32(const float* X, long long ld, int M, int N, float* out, int accumulate, float* scratch, void* stream);

/* ---- persistent LSTM recurrence (fp32 parity mode) ----------------------------------------------------------------
 * Replaces the time loop of nn.LSTM over a PackedSequence, pad_packed_sequence's zero fill and the locked-dropout
 * multiply: src/modules.py:78-84 (LockedLSTM) and :187-193 (pyramLockedLSTM); backward replaces autograd through them.
 *   gates   (B, T, ndir, 4H)  in: X.W_ih^T + b_ih + b_hh (gate order i,f,g,o) ; out: activated gates (saved)
 *   w_hh    (ndir, 4H, H)
 *   lens    (B) int32, 1 <= len <= T
 *   drop_mask (B, ndir*H) already scaled by 1/(1-p), or NULL
 *   out     (B, T, ndir*H) masked layer output, or NULL (then the caller uses hs_pad[:,1:T+1])
 *   hs_pad, cs_pad (B, T+2, ndir*H): frame t+1 = time t; frames 0 and T+1 are written as zeros
 * bwd: dout (B,T,ndir*H) contiguous; gates in: activated, out: d(pre-activation) (zeros at t >= len). */
size_t las_lstm_rec_workspace_bytes(int B, int H, int ndir);
int las_lstm_rec_fwd_f32(float* gates, const float* w_hh, const int* lens, const float* drop_mask, float* out, float* hs_pad,
                         float* cs_pad, int B, int T, int H, int ndir, void* ws, size_t ws_bytes, void* stream);
int las_lstm_rec_bwd_f32(const float* dout, float* gates, const float* cs_pad, const float* w_hh, const int* lens,
                         const float* drop_mask, int B, int T, int H, int ndir, void* ws, size_t ws_bytes, void* stream);

/* ---- persistent LSTM recurrence on the tensor pipe (bf16 operands, fp32 state) --------------------------------------
 * Same call sites and tensor contract as las_lstm_rec_fwd_f32, for the AMP path: W_hh is passed as bf16
 * (ndir, 4H, H); the 128 x H slice each CTA owns stays resident in shared memory as the A operand of tcgen05.mma and
 * h_{t-1} arrives by TMA each step.  H must be a multiple of 64; las_lstm_rec_tc_supported() says whether (B, H, ndir)
 * fits the co-resident grid.  save_gates = 0 skips writing the activated gates (inference). */
int las_lstm_rec_tc_supported(int B, int H, int ndir);
size_t las_lstm_rec_tc_workspace_bytes(int B, int H, int ndir);
int las_lstm_rec_fwd_tc(float* gates, const void* w_hh_bf16, const int* lens, const float* drop_mask, float* out, float* hs_pad,
                        float* cs_pad, int B, int T, int H, int ndir, int save_gates, void* ws, size_t ws_bytes, void* stream);
/* Same call sites (time loop of nn.LSTM on a PackedSequence + pad + locked dropout, src/modules.py:78-84,187-193), plus bf16
 * copies of the outputs that only GEMMs read afterwards (saves the separate fp32 -> bf16 cast passes):
 * out_bf16 (B, T, ndir*H) = the layer output after locked dropout (next layer's / key_map's / value_map's operand),
 * hs_bf16 (B, T+2, ndir*H) = the zero-framed hidden states (dW_hh operand of backward).  Either may be NULL; hs_pad may be
 * NULL when hs_bf16 is given, out when nobody reads the fp32 output. */
int las_lstm_rec_fwd_tc_ex(float* gates, const void* w_hh_bf16, const int* lens, const float* drop_mask, float* out, float* hs_pad,
                           float* cs_pad, int B, int T, int H, int ndir, int save_gates, void* ws, size_t ws_bytes,
                           void* out_bf16, void* hs_bf16, void* stream);

/* Arms the NEXT las_lstm_rec_bwd_tc / las_lstm_rec_bwd_tc_db call of this host thread: `side_stream` (a stream other than the one
 * the call is given; NULL disarms) is made to wait until every CTA of the recurrence kernel is resident -- the CTAs bump a device
 * counter on entry, the stream waits on its value -- or, on the variants without the counter, until the kernel has finished.
 * Work queued on side_stream afterwards can then fill the SMs the latency-bound recurrence leaves idle (the weight-gradient
 * GEMMs of the layer above, reference: autograd of src/modules.py:189) without delaying the recurrence's own cluster placement.
 * las_launch_start_mode(): 1 if the last armed call released side_stream at kernel start, 0 if at kernel end. */
void las_set_launch_start_stream(void* side_stream);
int las_launch_start_mode(void);

/* Forward counterpart: progress counters of the tensor-pipe forward recurrence, so that the NEXT layer's gate projection
 * (reference: the nn.LSTM input GEMM of src/modules.py:80 / :189 for layer l+1) can start on another stream, on the SMs the recurrence
 * leaves idle, for the time rows BOTH directions have already passed -- instead of after the whole layer.
 * las_lstm_rec_fwd_arm_progress(counters, every): the next las_lstm_rec_fwd_tc[_ex] call of this thread publishes: `counters` (>= 64
 *   zeroed 32-bit words, device memory, zeroed by the caller on the launch stream) gets one word per cluster (direction x batch-slice
 *   group); each CTA of the cluster adds 1 after every `every` steps once its outputs of those steps are visible device-wide.  Steps
 *   < k*every of a cluster are complete when its word >= k * ctas_per_cluster.
 * las_lstm_rec_fwd_progress_info(&clusters, &ctas_per_cluster): 1 if that launch publishes (clusters > 0); 0 if it ran a kernel that
 *   does not (the caller then waits for the launch to finish, as without the request).
 * las_stream_wait_value_geq(stream, word, value): cuStreamWaitValue32(GEQ) -- work queued on `stream` afterwards waits for the word. */
void las_lstm_rec_fwd_arm_progress(void* counters, int every);
/* same for the next las_lstm_rec_bwd_tc[_db] call: BPTT steps < k*every of a cluster have written their rows of the bf16 gate-gradient
 * matrix when its word >= k * ctas_per_cluster (the layer's dX GEMM then starts from the middle of the sequence outwards). */
void las_lstm_rec_bwd_arm_progress(void* counters, int every);
int las_lstm_rec_fwd_progress_info(int* clusters, int* ctas_per_cluster);
int las_stream_wait_value_geq(void* stream, const void* dev_word, unsigned value);
/* debug aid: device buffer (256*16 long long) receiving clock64 stamps of CTA (0,0,0) per timestep; NULL disables */
void las_lstm_rec_tc_set_debug(void* dev_buf);
/* BPTT on the tensor pipe.  w_hh_t_bf16 = W_hh transposed per direction, (ndir, H, 4H) bf16 (las_transpose_cast_bf16).
 * gates: in activated gates, out fp32 d(pre-activation); dgates_bf16 (B*T, ndir*4H): the same values as bf16, written
 * for every (b, t) (zeros at t >= len) -- the operand of the dX / dW_ih / dW_hh tensor-core GEMMs. */
int las_lstm_rec_bwd_tc(const float* dout, float* gates, void* dgates_bf16, const float* cs_pad, const void* w_hh_t_bf16,
                        const int* lens, const float* drop_mask, int B, int T, int H, int ndir, void* ws, size_t ws_bytes,
                        void* stream);
/* Same, with the bias gradients accumulated inside the kernel: dbias_partial (ndir, nslices, 4H) receives, per direction and
 * 32-row batch slice, sum over (rows of the slice, t) of d(pre-activation); the caller adds the nslices rows (fixed order).
 * In this form the fp32 write-back into `gates` is skipped (its only reader was the bias column sum); dgates_bf16 is
 * written as before.  las_lstm_rec_bwd_tc_dbias_slices returns nslices, or 0 when the form is not available for the shape. */
int las_lstm_rec_bwd_tc_dbias_slices(int B, int H, int ndir);
int las_lstm_rec_bwd_tc_db(const float* dout, float* gates, void* dgates_bf16, const float* cs_pad, const void* w_hh_t_bf16,
                           const int* lens, const float* drop_mask, int B, int T, int H, int ndir, void* ws, size_t ws_bytes,
                           float* dbias_partial, void* stream);
/* dst[b][c][r] (bf16) = src[b][r][c] (fp32) */
int las_transpose_cast_bf16(const float* src, void* dst, int batch, int rows, int cols, void* stream);

/* ---- fused attention step -------------------------------------------------------------------------------------------
 * Replaces MultiheadCrossAttention.forward after query_map (src/models.py:168-185): energy * sqrt(d), pad mask from
 * lengths (build_pad_masks :106-115, no host mask / H2D), softmax, zeroing, context.  K, V are (B, T, P) row-major.
 * fwd: q (row stride ld_q) -> ctx (ld_ctx) [+ ctx2 (ld_ctx2)], w ((B*heads), row stride ld_w) [+ w_b0: weights of
 *      batch row 0, (heads, T) contiguous -- the att_wgts[0] bookkeeping of src/models.py:349,377]
 * bwd: dctx (+ dctx2, nullable; when given the sum is written back to dctx) , w -> dq (ld_dq; accumulated into when
 *      dq_accumulate), de ((B*heads), ld_w) = d(energy) * scale, ready for the deferred dK = de^T.q GEMM.
 *      ctx (ld_ctx), when given, must be the context the forward call produced: sum_t w_t (dctx.V_t) == dctx.ctx lets
 *      backward run as ONE pass over K and V (T-split across a thread-block cluster); NULL selects the two-phase kernel.
 * Both directions read K and V exactly once: algorithmic bytes per call = 2 * B * T * P * sizeof(element). */
typedef struct {
    const float* q; long long ld_q;
    const float* K; const float* V; const int* lens;
    float* w; long long ld_w;
    float* w_b0;
    float* ctx; long long ld_ctx;
    float* ctx2; long long ld_ctx2;
    float* dctx; long long ld_dctx;
    const float* dctx2; long long ld_dctx2;
    float* dq; long long ld_dq; int dq_accumulate;
    float* de;
    int B, T, P, heads;
    float scale;
    void* ctx2_bf16; long long ld_ctx2_bf16;   /* fwd: optional bf16 copy of the context (next cell-0 GEMM operand) */
    void* dq_bf16; long long ld_dq_bf16;       /* bwd: optional bf16 copy of the total dq */
    int kv_bf16;                               /* 1: K and V point to bf16 (B,T,P) memory (AMP mode: half the bytes per step) */
    /* init-force prior (src/models.py:177-181): when fmask != NULL, ctx = softmax(w * fmask) . V over all T positions;
     * fmask row of (batch b, head h) = fmask + (b*heads+h)*ld_fmask (ld_fmask = 0: one (T) row shared by all).
     * w keeps the pre-prior weights (what the reference returns); w2 ((B*heads), ld_w) receives the second-softmax
     * weights in fwd and must be passed back to bwd. */
    const float* fmask; long long ld_fmask;
    float* w2;
    /* bwd: dctx2 may be the un-reduced output of a split-K GEMM: dctx2_nsplit (> 1) partial matrices, dctx2_split_stride
     * floats apart, are added up on the fly (0 / 1: a single matrix) */
    int dctx2_nsplit; long long dctx2_split_stride;
} LasAttnStep;
int las_attn_step_fwd_f32(const LasAttnStep* desc, void* stream);
int las_attn_step_bwd_f32(const LasAttnStep* desc, void* stream);

/* ---- single LSTM cell, pointwise part ---------------------------------------------------------------------------------
 * Replaces the fused pointwise of nn.LSTMCell + nn.Dropout in AutoRegDecoderLSTMCell.forward (src/modules.py:355-357)
 * for callers that step the cell themselves.  gates (B,4H): in pre-activations (both GEMMs + biases summed),
 * out activated gates.  bwd: gates in activated / out d(pre-activation); dc_io (B,H) carries dL/dc in and out. */
int las_lstm_cell_fwd_f32(float* gates, const float* c_prev, const float* mask, float* h_out, float* c_out, int B, int H,
                          void* stream);
int las_lstm_cell_bwd_f32(float* gates, const float* dh, const float* mask, const float* c, const float* c_prev, float* dc_io,
                          int B, int H, void* stream);

/* ---- Speller decoder loop --------------------------------------------------------------------------------------------
 * Replaces the Python loop of Speller.forward (src/models.py:336-385) including AutoRegDecoderLSTMCell.forward
 * (src/modules.py:340-365), the per-step attention, the tied classifier and the greedy argmax feedback, with no host
 * synchronisation per step (the reference syncs every step at src/models.py:377 and draws a host coin at :357 -- the
 * coins are pre-drawn by the caller into use_gold_host).
 * All matrices fp32 row-major.  `fws` / `iws` are caller-owned workspaces (las_speller_workspace_floats/ints); in
 * training they hold everything backward needs and must be passed unchanged to las_speller_bwd_f32. */
typedef struct {
    /* dims */
    int B, T, P, E, DH, DO, V, heads, steps;
    int sos_idx, pad_idx;
    int training;                 /* 1: save history for backward */
    int use_tc;                   /* 1: decoder GEMMs as bf16 tcgen05 tiles (AMP mode); fwd and bwd must agree */
    int kv_bf16;                  /* 1: K and V_ point to bf16 (B,T,P) memory; 2: IEEE fp16 (persistent decoder-step kernel only) */
    int init_force;               /* 1: block-diagonal attention prior on every loop step (src/models.py:326-330,364-366) */
    /* parameters */
    const float* emb;             /* (V, E)  char_emb.weight == cls.weight */
    const float* cls_b;           /* (V) */
    const float* w_ih0; const float* w_hh0; const float* b_ih0; const float* b_hh0;   /* (4DH, E+P) (4DH, DH) (4DH) */
    const float* w_ih1; const float* w_hh1; const float* b_ih1; const float* b_hh1;   /* (4DO, DH) (4DO, DO) (4DO) */
    const float* wq; const float* bq;   /* (P, DO) (P) */
    const float* init_query;      /* (DO) */
    /* inputs */
    const float* K; const float* V_; const int* enc_lens;   /* (B,T,P) (B,T,P) (B) */
    const int* dec_y; long long ld_y;   /* (B, >=steps) gold tokens, training only */
    const unsigned char* use_gold_host; /* HOST array (steps): step t>0 feeds the gold token y[:,t-1]; NULL => never */
    const float* drop0; const float* drop1;   /* (steps,B,DH) (steps,B,DO) scaled masks or NULL */
    /* outputs */
    float* logits;                /* (B, steps, V) */
    float* att0;                  /* (steps+1, heads, T): attention weights of batch row 0 */
    int* chars;                   /* (steps, B) greedy argmax per step (written when fed back or in eval) */
    /* workspaces */
    float* fws; size_t fws_floats;
    int* iws; size_t iws_ints;
    /* optional (backward, tensor-pipe mode): IEEE fp16 copies of K / V_ (B,T,P).  When both are given the backward loop's attention step
     * streams them (half the bytes) through mma.sync instead of the fp32 rows; NULL: fp32 rows */
    const void* K_f16; const void* V_f16;
} LasSpeller;
/* 1 when las_speller_fwd_f32 runs this shape as ONE persistent decoder-step kernel (tensor-pipe mode, heads == 1, no init_force,
 * dims that fit the TMEM-resident weight slices): K / V may then be fp32 (kv_bf16 = 0) or IEEE fp16 (kv_bf16 = 2), not bf16. */
int las_speller_persistent(int B, int T, int P, int DH, int DO, int V, int heads, int init_force, int use_tc);
size_t las_speller_workspace_floats(const LasSpeller* s);
size_t las_speller_workspace_ints(const LasSpeller* s);
int las_speller_fwd_f32(const LasSpeller* s, void* stream);
/* The forward loop is captured once per descriptor (all device pointers + the coin pattern) as a CUDA graph and replayed;
 * counts since library load: graphs captured / graph replays.  A steady loop must show replays only. */
void las_speller_graph_stats(long long* captures, long long* replays);

typedef struct {
    const float* dlogits;         /* (B, steps, V) contiguous */
    /* gradient outputs (overwritten) */
    float* d_emb; float* d_cls_b;
    float* d_w_ih0; float* d_w_hh0; float* d_b_ih0; float* d_b_hh0;
    float* d_w_ih1; float* d_w_hh1; float* d_b_ih1; float* d_b_hh1;
    float* d_wq; float* d_bq; float* d_init_query;
    float* dK; float* dV;         /* (B,T,P) */
} LasSpellerGrads;
int las_speller_bwd_f32(const LasSpeller* s, const LasSpellerGrads* g, void* stream);
/* The same backward in two separately enqueueable parts (autograd of src/models.py:336-385 has no such split: the reference computes
 * every gradient on one stream).  phases bit 0: the backward time loop, the initial attention step and dK / dV -- everything the
 * encoder's backward (src/models.py:60-66) waits for; bit 1: the batched parameter gradients (d_emb ... d_init_query), which read what
 * bit 0 left in the workspace and feed only the optimizer, so a caller may enqueue them on a second stream once the first part has
 * been enqueued (ordered behind it) and let them run beside the encoder's BPTT kernels.  phases = 3 is las_speller_bwd_f32. */
int las_speller_bwd_phases_f32(const LasSpeller* s, const LasSpellerGrads* g, int phases, void* stream);

/* ---- device-side collate + SpecAugment ---------------------------------------------------------------------------------
 * Replaces pad_sequence + FrequencyMasking / TimeMasking of datasetTrainDev.collate_fn (src/utils.py:95-128, maskers :82-84)
 * and the trainer's H2D copy of the padded batch (src/train.py:127).  frames: the length-sorted utterances back to back,
 * (sum(lens), F); offsets (B): first frame of each utterance; out (B, T, F) = frames padded with pad_value, then positions
 * with f in [f_lo, f_hi) or t in [t_lo, t_hi) set to mask_value (empty interval = no mask).  The mask intervals are drawn
 * by the host exactly like torchaudio.functional.mask_along_axis (one interval per axis for the whole batch). */
int las_collate_specaug_f32(const float* frames, const long long* offsets, const int* lens, int B, int T, int F, float pad_value,
                            int f_lo, int f_hi, int t_lo, int t_hi, float mask_value, float* out, void* stream);

/* ---- fused masked cross-entropy of the trainer --------------------------------------------------------------------------
 * Replaces the caller-side loss of src/train.py:117-136: y_mask = arange(L) < ly ; loss = sum(CE_none * y_mask) /
 * (n_nonpad * accu_grad) ; ppl = exp(loss) (:139), and the dev-eval variant that truncates a longer decode to the target
 * length (pred_logits[:, :L], :226-232).  logits: row (b, t) at logits + b*ld_b + t*V (ld_b = L*V for a contiguous (B, L, V)
 * tensor, steps*V for a truncated longer decode); y (B, >= L) int32 (row stride ld_y): the
 * targets AFTER the <sos> strip; ly_dev (B) int32 = ly - 1.  inv_denom = 1 / (n_nonpad * accu_grad), computed by the host
 * from the CPU length tensor (no device sync).  loss_ppl_out (2 floats): [loss, exp(loss)].  dlogits (nullable, (B*L, V)):
 * d loss / d logits.  Deterministic two-stage sum; scratch >= las_masked_ce_scratch_floats(B, L) floats. */
size_t las_masked_ce_scratch_floats(int B, int L);
int las_masked_ce_f32(const float* logits, long long ld_b, const int* y, long long ld_y, const int* ly_dev, int B, int L, int V, float inv_denom,
                      float* loss_ppl_out, float* dlogits, float* scratch, size_t scratch_floats, void* stream);

/* ---- fused unscale + global-norm clip + AdamW(amsgrad) ---------------------------------------------------------------
 * Replaces scaler.unscale_ -> clip_grad_norm_ -> scaler.step(AdamW amsgrad) (src/train.py:165-183; torch
 * optim/adam.py single-tensor math, nn/utils/clip_grad.py).  `table` is a device array of n_tensors LasAdamTensor;
 * `chunks` a device array of n_chunks {tensor, offset} pairs, `scratch` >= n_chunks + 8 floats.
 * Pass 1 computes sum((g*inv_scale)^2) and the non-finite flag; pass 2 applies coef = min(max_norm/(norm+1e-6), 1)
 * and the update, or skips everything when a non-finite gradient was seen (GradScaler.step semantics).
 * status (device, 2 floats): [0] = found_inf (0/1), [1] = total grad norm (unscaled). */
typedef struct {
    float* p; const float* g; float* m; float* v; float* vmax;
    long long numel;
    float step_size, bias_c2_sqrt;  /* lr/(1-beta1^step), sqrt(1-beta2^step): computed by the host in double */
} LasAdamTensor;
typedef struct { int tensor; int pad_; long long offset; } LasAdamChunk;
#define LAS_ADAM_CHUNK 65536
int las_adamw_amsgrad_fused(const LasAdamTensor* table, int n_tensors, const LasAdamChunk* chunks, int n_chunks, double lr,
                            double beta1, double beta2, float eps, double weight_decay, float inv_scale, float max_norm,
                            int amsgrad, float* scratch, float* status, void* stream);

/* ---- device-side transcript cut -------------------------------------------------------------------
 * Replaces the per-utterance host loop idx_to_str(pred_logits[b].argmax(-1), VOCAB, SOS_IDX, EOS_IDX) of src/infer.py:19-32,66
 * (and src/train.py:405-419): chars[t*ld_step + b*ld_b] (the greedy argmax, int32, device) -> out[b][0..lens[b]) = the token ids
 * with every <sos> dropped, cut before the first <eos>; out is (B, steps) bytes, lens (B) int32. */
int las_transcript_cut_i32(const int* chars, long long ld_step, long long ld_b, int B, int steps, int sos_idx, int eos_idx,
                           unsigned char* out, int* lens, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LAS_B200_H */
