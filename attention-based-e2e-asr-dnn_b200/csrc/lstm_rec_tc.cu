// Persistent LSTM recurrence on the tensor pipe (bf16 operands, fp32 accumulate / state), forward + BPTT.  sm_100a only.
//
// Same contract as lstm_rec_f32.cu (replaces the time loop of nn.LSTM over a PackedSequence, reference
// src/modules.py:78-82 / :187-191), different engine:
//   * CTA (r, s, dir) owns 32 hidden units (= 128 gate rows: 4 gates x 32 units) of direction `dir` for batch slice(s)
//     of 32 rows.  Its 128 x H slice of W_hh is loaded ONCE by TMA into shared memory (SWIZZLE_128B, K-major) and stays
//     resident for the whole sequence as the A operand of tcgen05.mma (M = 128 gate rows, N = 32 batch rows, K = H).
//   * per timestep: wait for the group's h_{t-1} (release/acquire counter), TMA-load the 32 x H bf16 slice of h_{t-1} as
//     the B operand, issue H/16 UMMAs into a TMEM accumulator, then the 4 epilogue warps (one per gate) read the
//     accumulator (tcgen05.ld), add the precomputed input projection, apply sigmoid / tanh, exchange the four gates
//     through shared memory, do the cell update with c_t held in registers, write h_t (bf16, for the next step's TMA),
//     the saved tensors and the (locked-dropout-masked) layer output, and publish the step.
//   * only the RS = H/32 CTAs that share (direction, batch slice) synchronise with one another -- the recurrence is
//     independent across batch rows -- so the barrier is 16-wide at H = 512, not grid-wide.
//   * a CTA may own several batch slices ("chains"): they are independent dependency chains, so while one chain waits
//     for its peers the tensor pipe and the epilogue work on the other.
#include "las_common.cuh"
#include "las_b200.h"
#include <cuda.h>
#include <mutex>
#include <stdlib.h>

static long long* g_rec_dbg = nullptr;   // debug aid, see las_lstm_rec_tc_set_debug

namespace {

constexpr int NB_SLICE = 32;          // batch rows per chain (UMMA N)
constexpr int UNITS = 32;             // hidden units per CTA
constexpr int ROWS = 4 * UNITS;       // gate rows per CTA (UMMA M)
constexpr int MAX_CHAINS = 2;
constexpr int NTHREADS = 256;

struct RecTcArgs {
    float* gates;          // (B, T, ndir, 4H) in: x-gates ; out: activated gates (when save_gates)
    const int* lens;
    const float* mask;     // (B, ndir*H) or null
    float* out;            // (B, T, ndir*H) or null
    float* hs_pad;         // (B, T+2, ndir*H) fp32
    float* cs_pad;         // (B, T+2, ndir*H) fp32
    __nv_bfloat16* out16;  // optional bf16 copy of `out` (B, T, ndir*H): the next layer's GEMM operand, written here instead of by a cast pass
    __nv_bfloat16* hs16;   // optional bf16 copy of hs_pad (B, T+2, ndir*H): the dW_hh GEMM operand of backward; hs_pad may then be null
    __nv_bfloat16* hbuf;   // (ndir, 2, Bpad, H) bf16 exchange buffer
    unsigned* ctr;         // (ndir, nslices) step counters
    int B, T, H, ndir, nslices, Bpad, chains, bsg, save;
    const __nv_bfloat16* w_gl;  // W_hh bf16 (ndir, 4H, H) for the TMEM-resident-A variant
    int w_tmem;            // 1: W_hh slice lives in TMEM (A operand from tensor memory); 0: in shared memory
    long long* dbg;        // optional (debug): per-step clock64 stamps of CTA (0,0,0), 16 slots per step
    unsigned* progress;    // optional (DSMEM kernel): one word per cluster (dir * gridDim.y + batch-slice group); every CTA adds 1 each time
    int progress_every;    //   its stores of another `progress_every` steps are visible device-wide (las_lstm_rec_fwd_arm_progress)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
                 "l"(map), "r"(bar), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
// A operand from tensor memory (".ts" form): D[tmem] (+)= A[tmem] . B[smem]
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
        "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]),
        "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ uint64_t make_desc_k(uint32_t saddr) {     // K-major, SWIZZLE_128B: LBO 16 B, SBO 1024 B
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// AMP-mode activations: one MUFU each (tanh.approx.f32, |err| ~ 2^-11, below bf16 operand rounding)
__device__ __forceinline__ float tanh_fast(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(0.5f, tanh_fast(0.5f * x), 0.5f); }
// one elected lane of a converged warp: unlike `if (lane == 0)` the compiler keeps the surrounding control flow warp-uniform, so
// the UMMA descriptors stay in uniform registers and each tcgen05.mma issues without a divergence-handling loop around it
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void cluster_arrive_release() { asm volatile("barrier.cluster.arrive.release;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait_acquire() { asm volatile("barrier.cluster.wait.acquire;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

#define REC_STAMP(slot)                                                                              \
    do {                                                                                            \
        if (a.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && s < 256) a.dbg[s * 16 + (slot)] = clock64(); \
    } while (0)

// idesc: F32 accumulate, BF16 x BF16, both K-major, M = 128, N = 32
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NB_SLICE >> 3) << 17) | ((uint32_t)(ROWS >> 4) << 24);

// the per-step operand tile (32 x H bf16) is fetched as NP pieces of KB/NP k-blocks, each with its own barrier, so the UMMAs of a
// piece run while the next pieces are still in flight
// measured at H = 512 (forward / BPTT us per timestep): 1 piece 3.78 / 4.96, 2 pieces 3.54 / 4.84, 4 pieces 3.86 / 5.15
__host__ __device__ inline int rec_pieces(int KB) { return (KB % 2 == 0) ? 2 : 1; }

constexpr int FWD_NACC = 4;          // independent TMEM accumulators per chain in the forward recurrence (one per k sub-step)

// recurrence kernels with eight epilogue warps: warps 0..3 as before (producer / helpers, UMMA issue), warps 4..11 epilogue
constexpr int NTHREADS8 = 384;
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

template <bool WTMEM>
__global__ void __launch_bounds__(NTHREADS8, 1) lstm_rec_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmW,
                                                                     const __grid_constant__ CUtensorMap tmH, const RecTcArgs a) {
    extern __shared__ uint8_t smem_raw[];
    const int H = a.H, T = a.T, KB = H / 64;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t w_sm = base;                                   // KB x [128 rows x 128 B] (absent when W lives in TMEM)
    const uint32_t h_sm = w_sm + (WTMEM ? 0 : KB * 16384);        // chains x KB x [32 rows x 128 B]
    const uint32_t ex_off = (h_sm - smem_u32(smem_raw)) + a.chains * KB * 4096;
    float* ex = reinterpret_cast<float*>(smem_raw + ex_off);      // [4][32][32] gate exchange
    const uint32_t bar_base = smem_u32(smem_raw) + ex_off + 4 * 32 * 32 * 4;
    auto full_bar = [&](int c) { return bar_base + 8u * c; };
    auto tfull_bar = [&](int c) { return bar_base + 8u * (MAX_CHAINS + c); };
    const uint32_t wbar = bar_base + 8u * (2 * MAX_CHAINS);
    const uint32_t tmem_slot = bar_base + 8u * (2 * MAX_CHAINS + 1);
    const int NP = rec_pieces(KB), KBP = KB / NP;
    auto piece_bar = [&](int c, int pc) { return pc == 0 ? full_bar(c) : bar_base + 8u * (2 * MAX_CHAINS + 2 + (pc - 1) * MAX_CHAINS + c); };
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int r = blockIdx.x, sg = blockIdx.y, dir = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int F = a.ndir * H;
    const long long brow = (long long)(T + 2) * F;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmH) : "memory");
        for (int c = 0; c < MAX_CHAINS; ++c) {
            mbar_init(tfull_bar(c), 1);
            for (int pc = 0; pc < 4; ++pc) mbar_init(piece_bar(c, pc), 1);
        }
        mbar_init(wbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // TMEM map: FWD_NACC independent accumulators per chain (MAX_CHAINS x FWD_NACC x 32 columns) first, then (WTMEM) the
    // resident W_hh slice
    const uint32_t tmem_cols = WTMEM ? 512u : 256u;
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    // TMEM map: accumulators (MAX_CHAINS x 32 columns) first, then the resident W_hh slice (H/2 columns: two bf16 per
    // 32-bit column, K ascending; TMEM lane = gate row, the layout tcgen05.mma expects for an A operand in tensor memory)
    const uint32_t tmem_w = tmem_base + MAX_CHAINS * FWD_NACC * NB_SLICE;
    if (WTMEM) {
        if (warp >= 4 && warp < 8) {
            const int q = warp & 3;
            const uint32_t* wrow = reinterpret_cast<const uint32_t*>(a.w_gl + ((long long)dir * 4 * H + q * H + r * UNITS + lane) * H);
            for (int cb = 0; cb < H / 64; ++cb) {
                uint32_t v[32];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const uint4 t4 = *reinterpret_cast<const uint4*>(wrow + cb * 32 + i * 4);
                    v[i * 4 + 0] = t4.x; v[i * 4 + 1] = t4.y; v[i * 4 + 2] = t4.z; v[i * 4 + 3] = t4.w;
                }
                tmem_st32(tmem_w + ((uint32_t)(q * 32) << 16) + cb * 32, v);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }

    if (warp == 0) {
        if (lane == 0) {
            // ===== producer: resident W_hh slice once (shared-memory variant), then h_{t-1} slices per step =====
            if (!WTMEM) {
                mbar_arrive_expect_tx(wbar, (uint32_t)KB * 16384u);
                for (int kb = 0; kb < KB; ++kb)
                    for (int g = 0; g < 4; ++g)
                        tma_load_2d(w_sm + kb * 16384 + g * 4096, &tmW, wbar, kb * 64, dir * 4 * H + g * H + r * UNITS);
            }
            for (int s = 1; s < T; ++s) {
                for (int c = 0; c < a.chains; ++c) {
                    const int slice = sg + c * a.bsg;
                    if (slice >= a.nslices) continue;
                    {
                        const unsigned* ctr = a.ctr + dir * a.nslices + slice;
                        const unsigned target = (unsigned)(H / UNITS) * (unsigned)s;
                        while (ld_acquire_gpu(ctr) < target) { }
                    }
                    REC_STAMP(0);
                    asm volatile("fence.proxy.async.global;" ::: "memory");       // peers' generic-proxy writes -> our TMA reads
                    REC_STAMP(11);
                    const int row0 = (dir * 2 + ((s - 1) & 1)) * a.Bpad + slice * NB_SLICE;
                    for (int pc = 0; pc < NP; ++pc) {                               // box (64, 32 rows, KBP k-blocks) per issue
                        mbar_arrive_expect_tx(piece_bar(c, pc), (uint32_t)KBP * 4096u);
                        tma_load_3d(h_sm + (c * KB + pc * KBP) * 4096, &tmH, piece_bar(c, pc), 0, row0, pc * KBP);
                    }
                    REC_STAMP(1);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: the whole warp walks the loop (waits included), one elected lane issues =====
        if (!WTMEM) mbar_wait(wbar, 0);
        for (int s = 1; s < T; ++s) {
            for (int c = 0; c < a.chains; ++c) {
                const int slice = sg + c * a.bsg;
                if (slice >= a.nslices) continue;
                for (int pc = 0; pc < NP; ++pc) {
                    mbar_wait(piece_bar(c, pc), (uint32_t)((s - 1) & 1));
                    if (lane == 0 && pc == 0) REC_STAMP(2);
                    tc_fence_after();
                    if (elect_one()) {
                        // consecutive UMMAs into ONE accumulator serialise on it: round-robin over FWD_NACC independent
                        // accumulators, summed by the epilogue
                        const int kb_lo = pc * KBP, kb_hi = kb_lo + KBP;
                        for (int kb = kb_lo; kb < kb_hi; ++kb) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const uint32_t d_tmem = tmem_base + (uint32_t)((c * FWD_NACC + k) * NB_SLICE);
                                const uint64_t bd = make_desc_k(h_sm + (c * KB + kb) * 4096 + k * 32);
                                if (WTMEM) {
                                    umma_bf16_ts(d_tmem, tmem_w + kb * 32 + k * 8, bd, IDESC, kb ? 1u : 0u);
                                } else {
                                    const uint64_t ad = make_desc_k(w_sm + kb * 16384 + k * 32);
                                    umma_bf16(d_tmem, ad, bd, IDESC, kb ? 1u : 0u);
                                }
                            }
                        }
                        if (kb_hi == KB) umma_commit(tfull_bar(c));
                    }
                    __syncwarp();
                }
                __syncwarp();
                if (lane == 0) REC_STAMP(3);
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: EIGHT warps (round 2): warp e = warp - 4 reads gate q = e & 3 (TMEM lane quarter) for batch columns
        // 16 (e >> 2) .. + 16 and updates the cells of batch rows 4e .. 4e + 3 =====
        const int e = warp - 4, q = e & 3, hf = e >> 2, j = lane;          // gate q, unit r*32 + j
        const int te = e * 32 + lane;              // 0..255
        const int u = r * UNITS + j;
        float cst[MAX_CHAINS][4];
        int lenr[MAX_CHAINS][4];      // lengths of this thread's 4 batch rows per chain (0 for rows >= B): loaded once
#pragma unroll
        for (int c = 0; c < MAX_CHAINS; ++c)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                cst[c][i] = 0.f;
                const int b = (sg + c * a.bsg) * NB_SLICE + e * 4 + i;
                lenr[c][i] = (c < a.chains && sg + c * a.bsg < a.nslices && b < a.B) ? a.lens[b] : 0;
            }
        // zero the pad frames (0 and T+1) of this CTA's (rows, units): the shifted h_{t-1} / c_{t-1} reads of backward
        for (int c = 0; c < a.chains; ++c) {
            const int slice = sg + c * a.bsg;
            if (slice >= a.nslices) continue;
            for (int i = 0; i < 4; ++i) {
                const int b = slice * NB_SLICE + e * 4 + i;
                if (b < a.B) {
                    const long long o0 = (long long)b * brow + dir * H + u, o1 = o0 + (long long)(T + 1) * F;
                    a.cs_pad[o0] = 0.f; a.cs_pad[o1] = 0.f;
                    if (a.hs_pad) { a.hs_pad[o0] = 0.f; a.hs_pad[o1] = 0.f; }
                    if (a.hs16) { a.hs16[o0] = __float2bfloat16(0.f); a.hs16[o1] = __float2bfloat16(0.f); }
                }
            }
        }
        for (int s = 0; s < T; ++s) {
            const int t = (dir == 0) ? s : (T - 1 - s);
#pragma unroll
            for (int c = 0; c < MAX_CHAINS; ++c) {
                const int slice = sg + c * a.bsg;
                if (c >= a.chains || slice >= a.nslices) continue;
                const int b0 = slice * NB_SLICE;
                // input projection for (gate q, unit j) of the 32 batch rows: coalesced 128 B per row, issued before the wait
                float xg[16];
                float* gbase = a.gates + ((long long)t * a.ndir + dir) * 4 * H + q * H + u;
                const long long gstride = (long long)T * a.ndir * 4 * H;
#pragma unroll
                for (int n = 0; n < 16; ++n) xg[n] = (b0 + 16 * hf + n < a.B) ? gbase[(long long)(b0 + 16 * hf + n) * gstride] : 0.f;
                if (te == 0) REC_STAMP(4);
                if (s > 0) {
                    mbar_wait(tfull_bar(c), (uint32_t)((s - 1) & 1));
                    if (te == 0) REC_STAMP(5);
                    tc_fence_after();
#pragma unroll
                    for (int acc = 0; acc < FWD_NACC; ++acc) {
                        uint32_t v[16];
                        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((c * FWD_NACC + acc) * NB_SLICE + 16 * hf), v);
#pragma unroll
                        for (int n = 0; n < 16; ++n) xg[n] += __uint_as_float(v[n]);
                    }
                    if (te == 0) REC_STAMP(6);
                }
#pragma unroll
                for (int n = 0; n < 16; ++n) {
                    const float pre = xg[n];
                    const float act = (q == 2) ? tanh_fast(pre) : sigmoid_fast(pre);
                    ex[(q * 32 + 16 * hf + n) * 32 + j] = act;
                    xg[n] = act;                                   // kept for the deferred save below
                }
                tc_fence_before();
                named_bar_sync(1, 256);
                if (te == 0) REC_STAMP(7);
                // cell update: thread (e, j) owns unit j for batch rows n = 4e + i
                float hh[4], cc[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int n = e * 4 + i, b = b0 + n;
                    const float gi = ex[(0 * 32 + n) * 32 + j], gf = ex[(1 * 32 + n) * 32 + j];
                    const float gg = ex[(2 * 32 + n) * 32 + j], go = ex[(3 * 32 + n) * 32 + j];
                    const bool valid = t < lenr[c][i];
                    cc[i] = 0.f; hh[i] = 0.f;
                    if (valid) {
                        cc[i] = fmaf(gf, cst[c][i], gi * gg);
                        hh[i] = go * tanh_fast(cc[i]);
                    }
                    cst[c][i] = cc[i];
                    // the only data peers wait for: h_t as bf16
                    a.hbuf[((long long)(dir * 2 + (s & 1)) * a.Bpad + b) * H + u] = __float2bfloat16(hh[i]);
                }
                // publish step s of this chain FIRST (the release only has the 4 bf16 stores per thread in front of it) ...
                named_bar_sync(1, 256);
                if (te == 0) REC_STAMP(8);
                if (te == 0 && s + 1 < T) red_release_gpu_add(a.ctr + dir * a.nslices + slice, 1u);
                if (te == 0) REC_STAMP(9);
                // ... then write what only backward / the next layer read; these stores overlap the wait for the next step
                if (a.save) {
#pragma unroll
                    for (int n = 0; n < 16; ++n)
                        if (b0 + 16 * hf + n < a.B) gbase[(long long)(b0 + 16 * hf + n) * gstride] = xg[n];
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int b = b0 + e * 4 + i;
                    if (b < a.B) {
                        const long long so = (long long)b * brow + (long long)(t + 1) * F + dir * H + u;
                        if (a.hs_pad) a.hs_pad[so] = hh[i];
                        if (a.hs16) a.hs16[so] = __float2bfloat16(hh[i]);
                        a.cs_pad[so] = cc[i];
                        if (a.out || a.out16) {
                            const float m = a.mask ? a.mask[(long long)b * F + dir * H + u] : 1.f;
                            const long long oo = ((long long)b * T + t) * F + dir * H + u;
                            if (a.out) a.out[oo] = hh[i] * m;
                            if (a.out16) a.out16[oo] = __float2bfloat16(hh[i] * m);
                        }
                    }
                }
                if (te == 0) REC_STAMP(10);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}


// =====================================================================================================================
// DSMEM exchange variant of the forward recurrence (the default when the group fits one thread-block cluster).
// The RS = H/32 CTAs of a (direction, batch slice) group form ONE cluster.  h_t never goes through global memory / L2 on
// its way to the peers: each CTA stages its 32 rows x 32 units bf16 slice (2 KB) in shared memory in the UMMA K-major
// *no-swizzle* core-matrix layout and pushes it into every peer's operand buffer with one bulk async copy per peer
// (cp.async.bulk.shared::cluster.shared::cta), which also signals the peer's mbarrier (complete_tx).  No counter, no
// release fence waiting for L2 acks, no polling, no TMA fetch: the MMA warp just waits for 2 x RS/2 x 2 KB to land.
// Measured (scripts/micro/dsmem_bulk.cu): the 16-way all-to-all of 2 KB slices completes ~1950 cycles after the first issue
// (~17 B/clk into each SM) against ~3750 for release -> counter -> poll -> proxy fence -> TMA.
//   * operand layout: core matrix = 8 batch rows x 8 units (128 B); address(kc, ng) = kc*512 + ng*128 with kc = unit/8,
//     ng = row/8 -> UMMA descriptor SWIZZLE_NONE, LBO (next core along K) = 512 B, SBO (next 8 rows) = 128 B; the slice of
//     source CTA r is the contiguous range [r*2048, (r+1)*2048).
//   * two operand buffers (step parity): a peer can be at most one step ahead of this CTA's tensor pipe.
//   * sources 0..RS/2-1 signal barrier A, the others barrier B: the UMMAs of the first half of K start while the second
//     half is still arriving.
// =====================================================================================================================
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void bulk_copy_to_peer(uint32_t dst_cluster_addr, uint32_t src_cta_addr, uint32_t bytes, uint32_t remote_bar) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_cluster_addr),
                 "r"(src_cta_addr), "r"(bytes), "r"(remote_bar)
                 : "memory");
}
__device__ __forceinline__ uint64_t make_desc_k_noswz(uint32_t saddr) {     // K-major, no swizzle: LBO 512 B, SBO 128 B
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)(512 >> 4) << 16;
    d |= (uint64_t)(128 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// =====================================================================================================================
// The kernel: 384 threads = warps 0, 2, 3 helpers, warp 1 UMMA issue, warps 4..11 epilogue.
//   * epilogue warps 4+q and 8+q share TMEM lane quarter q (= gate q) and split the 32 batch columns of the tile in halves: every
//     thread activates 16 gates and updates 4 cells (round 2; with four epilogue warps -- 32 activations and 8 cells per thread -- a
//     step took 2.45 us instead of 2.34);
//   * the helper warps do everything in global memory that is not the step-to-step dependency: they stage the input projection
//     x-gates of step s+1 in shared memory while step s runs and copy the activated gates of step s out (backward reads them); the
//     epilogue warps keep only the h / c / out stores: after publishing, an SM's global-memory instructions queue behind its
//     outgoing DSMEM copies, and that queue -- not the exchange -- had become the critical path;
//   * optional progress counters for consumers on other streams (las_lstm_rec_fwd_arm_progress).
// One chain per CTA.
// =====================================================================================================================

__global__ void __launch_bounds__(NTHREADS8, 1) lstm_rec_fwd_dsm_kernel(const RecTcArgs a) {
    extern __shared__ uint8_t smem_raw[];
    const int H = a.H, T = a.T;
    const int RS = H / UNITS;                                     // CTAs per group = cluster size
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t TILE = (uint32_t)RS * 2048u;                   // 32 rows x H bf16
    const uint32_t h_sm = base;                                   // [parity][TILE]
    const uint32_t stage_sm = h_sm + 2 * TILE;                    // 2 KB: this CTA's slice in core-matrix layout
    const uint32_t ex_off = (stage_sm - smem_u32(smem_raw)) + 2048;
    float* ex = reinterpret_cast<float*>(smem_raw + ex_off);      // [4][32][32] gate exchange
    float* xgs = ex + 4 * 32 * 32;                                // [2][4][32][32] input projection of the next steps
    const uint32_t bar_base = smem_u32(smem_raw) + ex_off + 3 * 4 * 32 * 32 * 4;
    auto hbar = [&](int par, int half) { return bar_base + 8u * (par * 2 + half); };
    const uint32_t tfull_bar = bar_base + 8u * 8;
    const uint32_t tmem_slot = bar_base + 8u * 10;
    const uint32_t tempty_bar = bar_base + 8u * 11;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
    uint8_t* stage_ptr = smem_raw + (stage_sm - smem_u32(smem_raw));

    const int r = blockIdx.x, sg = blockIdx.y, dir = blockIdx.z;   // r == rank in the cluster (cluster dims (RS,1,1))
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int F = a.ndir * H;
    const long long brow = (long long)(T + 2) * F;
    const int halfsrc = RS >= 2 ? RS / 2 : 1;                     // sources [0, halfsrc) -> barrier A, the rest -> barrier B
    const int nhalf = RS >= 2 ? 2 : 1;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < 4; ++i) mbar_init(bar_base + 8u * i, 1);
        mbar_init(tfull_bar, 1);
        mbar_init(tempty_bar, 8);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const uint32_t tmem_cols = 512u;
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    const uint32_t tmem_w = tmem_base + MAX_CHAINS * FWD_NACC * NB_SLICE;
    if (warp >= 4 && warp < 8) {
        // resident A operand: this CTA's 128 x H slice of W_hh in tensor memory (TMEM lane = gate row, two bf16 per column)
        const int qq = warp & 3;
        const uint32_t* wrow = reinterpret_cast<const uint32_t*>(a.w_gl + ((long long)dir * 4 * H + qq * H + r * UNITS + lane) * H);
        for (int cb = 0; cb < H / 64; ++cb) {
            uint32_t v[32];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint4 t4 = *reinterpret_cast<const uint4*>(wrow + cb * 32 + i * 4);
                v[i * 4 + 0] = t4.x; v[i * 4 + 1] = t4.y; v[i * 4 + 2] = t4.z; v[i * 4 + 3] = t4.w;
            }
            tmem_st32(tmem_w + ((uint32_t)(qq * 32) << 16) + cb * 32, v);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // every CTA of the cluster has initialised its barriers before anyone pushes data at it
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");

    if (warp == 1) {
        // ===== MMA issuer (whole warp walks the loop, one elected lane issues) =====
        for (int s = 1; s < T; ++s) {
            if (sg >= a.nslices) continue;
            const int par = (s - 1) & 1;                          // buffer holding h_{s-1}
            const uint32_t phase = (uint32_t)(((s - 1) >> 1) & 1);
            const uint32_t buf = h_sm + (uint32_t)par * TILE;
            if (s > 1) mbar_wait(tempty_bar, (uint32_t)(s & 1));          // the epilogue has read step s-1's accumulators
            for (int hf = 0; hf < nhalf; ++hf) {
                if (lane == 0) mbar_arrive_expect_tx(hbar(par, hf), (uint32_t)(hf == 0 ? halfsrc : RS - halfsrc) * 2048u);
                __syncwarp();
                mbar_wait(hbar(par, hf), phase);
                if (lane == 0 && hf == 0) REC_STAMP(2);
                tc_fence_after();
                if (elect_one()) {
                    const int ks_lo = hf == 0 ? 0 : halfsrc * 2, ks_hi = hf == 0 ? (nhalf == 2 ? halfsrc * 2 : RS * 2) : RS * 2;    // 16-unit k-steps
                    for (int ks = ks_lo; ks < ks_hi; ++ks) {
                        const uint32_t d_tmem = tmem_base + (uint32_t)((ks & 3) * NB_SLICE);
                        const uint64_t bd = make_desc_k_noswz(buf + (uint32_t)ks * 1024u);
                        umma_bf16_ts(d_tmem, tmem_w + ks * 8, bd, IDESC, ks >= 4 ? 1u : 0u);
                    }
                    if (ks_hi == RS * 2) umma_commit(tfull_bar);
                }
                __syncwarp();
            }
            if (lane == 0) REC_STAMP(3);
        }
    } else if (warp < 4) {
        // ===== helper warps (0, 2, 3): x-gate staging for the next step, gate saves of this one =====
        if (sg < a.nslices) {
            const int hw = warp == 0 ? 0 : warp - 1;       // 0..2
            const int b0 = sg * NB_SLICE;
            const long long gstride = (long long)T * a.ndir * 4 * H;
            const int rg = lane >> 3, c4 = 4 * (lane & 7);
            float4 xr[11];
            auto load_x = [&](int s2) {
                const int t2 = (dir == 0) ? s2 : (T - 1 - s2);
                const float* gb = a.gates + ((long long)t2 * a.ndir + dir) * 4 * H + r * UNITS + c4;
#pragma unroll
                for (int i = 0; i < 11; ++i) {
                    const int pr = 4 * (hw + 3 * i) + rg;          // pr = gate * 32 + batch row
                    const int qq = pr >> 5, n = pr & 31;
                    xr[i] = (pr < 128 && b0 + n < a.B) ? *reinterpret_cast<const float4*>(gb + (long long)(b0 + n) * gstride + qq * H)
                                                       : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            };
            auto store_x = [&](int buf) {
#pragma unroll
                for (int i = 0; i < 11; ++i) {
                    const int pr = 4 * (hw + 3 * i) + rg;
                    if (pr < 128) *reinterpret_cast<float4*>(xgs + (buf * 128 + pr) * 32 + c4) = xr[i];
                }
            };
            load_x(0);
            store_x(0);
            asm volatile("bar.arrive 4, 352;" ::: "memory");                      // x-gates of step 0 staged
            for (int s = 0; s < T; ++s) {
                const int t = (dir == 0) ? s : (T - 1 - s);
                if (s + 1 < T) load_x(s + 1);                                      // in flight while waiting below
                asm volatile("bar.sync 2, 352;" ::: "memory");                    // `ex` of step s complete (and xgs[s & 1] consumed)
                // the epilogue warps stored h / c / out of step s-1 before arriving at that barrier: every `progress_every` steps one
                // helper thread makes them visible device-wide and counts this CTA in (a consumer on another stream waits for the
                // cluster's count: the next layer's gate GEMM starts on the rows both directions have passed)
                if (a.progress && warp == 0 && lane == 0 && s > 0 && s % a.progress_every == 0) {
                    __threadfence();
                    atomicAdd(a.progress + dir * gridDim.y + sg, 1u);
                }
                if (a.save) {
                    float* gb = a.gates + ((long long)t * a.ndir + dir) * 4 * H + r * UNITS + c4;
                    for (int grp = hw; grp < 32; grp += 3) {
                        const int pr = 4 * grp + rg;
                        const int qq = pr >> 5, n = pr & 31;
                        if (b0 + n < a.B)
                            *reinterpret_cast<float4*>(gb + (long long)(b0 + n) * gstride + qq * H) = *reinterpret_cast<const float4*>(ex + pr * 32 + c4);
                    }
                }
                asm volatile("bar.arrive 3, 352;" ::: "memory");                  // `ex` may be rewritten
                if (s + 1 < T) {
                    store_x((s + 1) & 1);                                          // that buffer was consumed at step s-1
                    asm volatile("bar.arrive 4, 352;" ::: "memory");              // x-gates of step s+1 staged
                }
            }
        }
    } else if (sg < a.nslices) {
        // ===== epilogue: warp e = warp - 4: gate q = e & 3 (TMEM lane quarter), column half hf = e >> 2 =====
        const int e = warp - 4, q = e & 3, hf = e >> 2, j = lane;
        const int te = e * 32 + lane;              // 0..255
        const int u = r * UNITS + j;
        const int b0 = sg * NB_SLICE;
        float xg[16];
        float cst[4], mreg[4];
        int lenr[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            cst[i] = 0.f;
            const int b = b0 + e * 4 + i;           // cell phase: warp e owns batch rows 4e .. 4e+3
            lenr[i] = b < a.B ? a.lens[b] : 0;
            mreg[i] = (a.mask && b < a.B) ? a.mask[(long long)b * F + dir * H + u] : 1.f;
            if (b < a.B) {
                const long long o0 = (long long)b * brow + dir * H + u, o1 = o0 + (long long)(T + 1) * F;
                a.cs_pad[o0] = 0.f; a.cs_pad[o1] = 0.f;
                if (a.hs_pad) { a.hs_pad[o0] = 0.f; a.hs_pad[o1] = 0.f; }
                if (a.hs16) { a.hs16[o0] = __float2bfloat16(0.f); a.hs16[o1] = __float2bfloat16(0.f); }
            }
        }
        for (int s = 0; s < T; ++s) {
            const int t = (dir == 0) ? s : (T - 1 - s);
            asm volatile("bar.sync 4, 352;" ::: "memory");              // the helper warps have staged this step's x-gates
#pragma unroll
            for (int n = 0; n < 16; ++n) xg[n] = xgs[((s & 1) * 128 + q * 32 + 16 * hf + n) * 32 + j];
            if (te == 0) REC_STAMP(4);
            if (s > 0) {
                mbar_wait(tfull_bar, (uint32_t)((s - 1) & 1));
                if (te == 0) REC_STAMP(5);
                tc_fence_after();
#pragma unroll
                for (int acc = 0; acc < FWD_NACC; ++acc) {
                    uint32_t v[16];
                    tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * NB_SLICE + 16 * hf), v);
#pragma unroll
                    for (int n = 0; n < 16; ++n) xg[n] += __uint_as_float(v[n]);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty_bar);
                if (te == 0) REC_STAMP(6);
            }
            if (s > 0) asm volatile("bar.sync 3, 352;" ::: "memory");       // the savers are done with the previous step's `ex`
#pragma unroll
            for (int n = 0; n < 16; ++n) ex[(q * 32 + 16 * hf + n) * 32 + j] = (q == 2) ? tanh_fast(xg[n]) : sigmoid_fast(xg[n]);
            tc_fence_before();
            // the bulk copies of the previous step must have finished READING the staging tile before it is rewritten
            if (te < RS && s > 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            named_bar_sync(1, 256);
            asm volatile("bar.arrive 2, 352;" ::: "memory");               // `ex` complete: the savers may copy it out
            if (te == 0) REC_STAMP(7);
            float hh[4], cc[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int n = e * 4 + i;
                const float gi = ex[(0 * 32 + n) * 32 + j], gf = ex[(1 * 32 + n) * 32 + j];
                const float gg = ex[(2 * 32 + n) * 32 + j], go = ex[(3 * 32 + n) * 32 + j];
                const bool valid = t < lenr[i];
                cc[i] = 0.f; hh[i] = 0.f;
                if (valid) {
                    cc[i] = fmaf(gf, cst[i], gi * gg);
                    hh[i] = go * tanh_fast(cc[i]);
                }
                cst[i] = cc[i];
            }
            if (s + 1 < T) {
                // stage h_t (bf16) in core-matrix layout: core (kc = j/8, ng = row/8 = e>>1), row-in-core (e&1)*4 + i, element j%8
                __nv_bfloat16* st = reinterpret_cast<__nv_bfloat16*>(stage_ptr + (j >> 3) * 512 + (e >> 1) * 128 + (e & 1) * 64) + (j & 7);
#pragma unroll
                for (int i = 0; i < 4; ++i) st[i * 8] = __float2bfloat16(hh[i]);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            }
            named_bar_sync(1, 256);
            if (te == 0) REC_STAMP(8);
            if (s + 1 < T && te < RS) {
                // thread `te` pushes this CTA's slice into peer te's buffer for step s+1 (parity s&1) and signals its barrier
                const int par = s & 1;
                const uint32_t dst = mapa_u32(h_sm + (uint32_t)par * TILE + (uint32_t)r * 2048u, (uint32_t)te);
                const uint32_t rbar = mapa_u32(hbar(par, (nhalf == 2 && r >= halfsrc) ? 1 : 0), (uint32_t)te);
                bulk_copy_to_peer(dst, stage_sm, 2048u, rbar);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            if (te == 0) REC_STAMP(9);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int b = b0 + e * 4 + i;
                if (b < a.B) {
                    const long long so = (long long)b * brow + (long long)(t + 1) * F + dir * H + u;
                    if (a.hs_pad) a.hs_pad[so] = hh[i];
                    if (a.hs16) a.hs16[so] = __float2bfloat16(hh[i]);
                    a.cs_pad[so] = cc[i];
                    if (a.out || a.out16) {
                        const float m = mreg[i];
                        const long long oo = ((long long)b * T + t) * F + dir * H + u;
                        if (a.out) a.out[oo] = hh[i] * m;
                        if (a.out16) a.out16[oo] = __float2bfloat16(hh[i] * m);
                    }
                }
            }
            if (te == 0) REC_STAMP(10);
        }
        if (te < RS) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    // no CTA of the cluster may exit while a peer's copy into its shared memory could still be in flight
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode2() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = (EncodeTiledFn)p;
    return fn;
}
int make_map_2d(CUtensorMap* m, const void* ptr, long long cols, long long rows, int box_cols, int box_rows) {
    EncodeTiledFn enc = get_encode2();
    if (!enc) { las_set_error("cuTensorMapEncodeTiled entry point not available"); return LAS_ERR_CUDA; }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult rc = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) { las_set_error("cuTensorMapEncodeTiled (2d) failed (%d)", (int)rc); return LAS_ERR_CUDA; }
    return LAS_OK;
}

// rank-N bf16 map: dims[0] contiguous; strides_elems[i] = stride of dims[i+1] in elements
int make_map_nd(CUtensorMap* m, const void* ptr, int rank, const long long* dims_, const long long* strides_elems, const int* box_) {
    EncodeTiledFn enc = get_encode2();
    if (!enc) { las_set_error("cuTensorMapEncodeTiled entry point not available"); return LAS_ERR_CUDA; }
    cuuint64_t dims[5], strides[4];
    cuuint32_t box[5], estr[5];
    for (int i = 0; i < rank; ++i) { dims[i] = (cuuint64_t)dims_[i]; box[i] = (cuuint32_t)box_[i]; estr[i] = 1; }
    for (int i = 0; i + 1 < rank; ++i) strides[i] = (cuuint64_t)strides_elems[i] * 2;
    CUresult rc = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) { las_set_error("cuTensorMapEncodeTiled (rank %d) failed (%d)", rank, (int)rc); return LAS_ERR_CUDA; }
    return LAS_OK;
}

struct Plan { int rs, nslices, bsg, chains, Bpad; size_t smem; };
int make_plan(int B, int H, int ndir, Plan* p) {
    LAS_CHECK_ARG(H % 64 == 0 && H >= 64, "lstm_rec_tc: hidden size %d must be a multiple of 64", H);
    const LasDeviceInfo* di = las_device_info();
    p->rs = H / UNITS;
    p->nslices = ceil_div(B, NB_SLICE);
    int max_bsg = di->num_sms / (p->rs * ndir);
    LAS_CHECK_ARG(max_bsg >= 1, "lstm_rec_tc: H=%d needs more co-resident CTAs than SMs", H);
    p->bsg = p->nslices < max_bsg ? p->nslices : max_bsg;
    p->chains = ceil_div(p->nslices, p->bsg);
    LAS_CHECK_ARG(p->chains <= MAX_CHAINS, "lstm_rec_tc: batch %d needs %d chains per CTA (max %d) at H=%d", B, p->chains, MAX_CHAINS, H);
    p->Bpad = p->nslices * NB_SLICE;
    const int KB = H / 64;
    p->smem = 1024 + (size_t)KB * 16384 + (size_t)p->chains * KB * 4096 + 4 * 32 * 32 * 4 + 8 * (2 * MAX_CHAINS + 2) + 64;
    LAS_CHECK_ARG(p->smem <= (size_t)di->max_smem_optin, "lstm_rec_tc: needs %zu B of shared memory", p->smem);
    return LAS_OK;
}

}  // namespace

// debug aid: device buffer of 256*16 long long receiving clock64 stamps of CTA (0,0,0) of the next forward launches
extern "C" void las_lstm_rec_tc_set_debug(void* dev_buf) { g_rec_dbg = (long long*)dev_buf; }

extern "C" int las_lstm_rec_tc_supported(int B, int H, int ndir) {
    Plan p;
    if (B < 1 || (ndir != 1 && ndir != 2)) return 0;
    if (H % 64 != 0 || H < 64) return 0;
    const LasDeviceInfo* di = las_device_info();
    int rs = H / UNITS, nsl = ceil_div(B, NB_SLICE);
    int max_bsg = di->num_sms / (rs * ndir);
    if (max_bsg < 1) return 0;
    int bsg = nsl < max_bsg ? nsl : max_bsg;
    if (ceil_div(nsl, bsg) > MAX_CHAINS) return 0;
    (void)p;
    return 1;
}

extern "C" size_t las_lstm_rec_tc_workspace_bytes(int B, int H, int ndir) {
    const int nsl = ceil_div(B, NB_SLICE);
    // counters (1 KB) + exchange buffer of tagged 8-byte words {2 x bf16, step}: forward h (ndir, 2, Bpad, H/2);
    // backward dG (ndir, 2, 4, Bpad, H/2)
    return 1024 + (size_t)ndir * 2 * 4 * nsl * NB_SLICE * H * 4;
}

// ---- forward progress counters (next layer's gate GEMM pipelined behind this layer's recurrence) ----
static thread_local unsigned* t_prog_ctr = nullptr;
static thread_local int t_prog_every = 0, t_prog_clusters = 0, t_prog_rs = 0;
extern "C" void las_lstm_rec_fwd_arm_progress(void* counters, int every) {
    t_prog_ctr = (unsigned*)counters;
    t_prog_every = every;
    t_prog_clusters = 0;
    t_prog_rs = 0;
}
// backward: the BPTT kernel publishes the same way (steps < k * every of a cluster have their d(pre-activation) rows in the bf16
// gate-gradient matrix when its word >= k * ctas_per_cluster); las_lstm_rec_fwd_progress_info reports the last armed launch of either kind
static thread_local unsigned* t_bprog_ctr = nullptr;
static thread_local int t_bprog_every = 0;
extern "C" void las_lstm_rec_bwd_arm_progress(void* counters, int every) {
    t_bprog_ctr = (unsigned*)counters;
    t_bprog_every = every;
    t_prog_clusters = 0;
    t_prog_rs = 0;
}
extern "C" int las_lstm_rec_fwd_progress_info(int* clusters, int* ctas_per_cluster) {
    if (clusters) *clusters = t_prog_clusters;
    if (ctas_per_cluster) *ctas_per_cluster = t_prog_rs;
    return t_prog_clusters > 0 ? 1 : 0;
}
typedef CUresult (*WaitValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
static WaitValue32Fn get_wait_value32() {
    // (driver entry point looked up at run time: the library must load on machines without libcuda, e.g. for the CPU-side ABI tests)
    static WaitValue32Fn fn = nullptr;
    if (!fn) {
        void* pfn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &pfn, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = (WaitValue32Fn)pfn;
        else
            cudaGetLastError();
    }
    return fn;
}
extern "C" int las_stream_wait_value_geq(void* stream, const void* dev_word, unsigned value) {
    WaitValue32Fn fn = get_wait_value32();
    if (!fn) { las_set_error("cuStreamWaitValue32 is not available"); return LAS_ERR_UNSUPPORTED; }
    const CUresult rc = fn((CUstream)stream, (CUdeviceptr)dev_word, value, CU_STREAM_WAIT_VALUE_GEQ);
    if (rc != CUDA_SUCCESS) { las_set_error("cuStreamWaitValue32 failed (%d)", (int)rc); return LAS_ERR_CUDA; }
    return LAS_OK;
}

static int rec_fwd_tc_impl(float* gates, const void* w_hh_bf16, const int* lens, const float* drop_mask, float* out, float* hs_pad,
                           float* cs_pad, int B, int T, int H, int ndir, int save_gates, void* ws, size_t ws_bytes,
                           __nv_bfloat16* out16, __nv_bfloat16* hs16, void* stream) {
    LAS_CHECK_ARG(gates && w_hh_bf16 && lens && (hs_pad || hs16) && cs_pad && ws, "lstm_rec_fwd_tc: null pointer");
    LAS_CHECK_ARG(B >= 1 && T >= 1 && (ndir == 1 || ndir == 2), "lstm_rec_fwd_tc: bad dims");
    int rc = las_set_device_of(gates);
    if (rc) return rc;
    Plan p;
    rc = make_plan(B, H, ndir, &p);
    if (rc) return rc;
    if (ws_bytes < las_lstm_rec_tc_workspace_bytes(B, H, ndir)) { las_set_error("lstm_rec_fwd_tc: workspace too small"); return LAS_ERR_WORKSPACE; }
    cudaStream_t st = (cudaStream_t)stream;
    // a progress request is for THIS launch only (one chain per CTA, DSMEM kernel); anything else leaves progress_info() at 0 clusters
    unsigned* prog_ctr = t_prog_ctr;
    const int prog_every = t_prog_every;
    t_prog_ctr = nullptr; t_prog_every = 0; t_prog_clusters = 0; t_prog_rs = 0;
    {
        // LAS_REC_SPLIT_BATCH=1: with more batch slices than one chain per CTA allows (B > 128 at H = 512), run the rows in passes of
        // `bsg` slices with the one-chain DSMEM kernel (every array is batch-major: a pass is a pointer offset).  Measured SLOWER at
        // B = 256: a pass of 128 rows is 8 clusters of 16 CTAs, and with all 8 GPCs exchanging at once a step takes 4.9 us (2.5 us
        // with 6 clusters) -- 9.7 us per step for both passes against 6.9 us for two chains per CTA on the counter/TMA kernel.
        const char* sb = getenv("LAS_REC_SPLIT_BATCH");
        const char* de = getenv("LAS_REC_DSMEM");
        if (p.chains > 1 && (sb && atoi(sb) == 1) && !(de && atoi(de) == 0) && p.rs <= 16 && H <= 512) {
            const int rows = p.bsg * NB_SLICE;
            const long long F = (long long)ndir * H;
            for (int b0 = 0; b0 < B; b0 += rows) {
                const int nb = B - b0 < rows ? B - b0 : rows;
                rc = rec_fwd_tc_impl(gates + (long long)b0 * T * ndir * 4 * H, w_hh_bf16, lens + b0, drop_mask ? drop_mask + b0 * F : nullptr,
                                     out ? out + (long long)b0 * T * F : nullptr, hs_pad ? hs_pad + (long long)b0 * (T + 2) * F : nullptr,
                                     cs_pad + (long long)b0 * (T + 2) * F, nb, T, H, ndir, save_gates, ws, ws_bytes,
                                     out16 ? out16 + (long long)b0 * T * F : nullptr, hs16 ? hs16 + (long long)b0 * (T + 2) * F : nullptr, stream);
                if (rc) return rc;
            }
            return LAS_OK;
        }
    }
    RecTcArgs a{};
    a.gates = gates; a.lens = lens; a.mask = drop_mask; a.out = out; a.hs_pad = hs_pad; a.cs_pad = cs_pad; a.out16 = out16; a.hs16 = hs16;
    a.ctr = (unsigned*)ws; a.hbuf = (__nv_bfloat16*)((char*)ws + 1024);
    a.B = B; a.T = T; a.H = H; a.ndir = ndir; a.nslices = p.nslices; a.Bpad = p.Bpad; a.chains = p.chains; a.bsg = p.bsg; a.save = save_gates; a.dbg = g_rec_dbg;
    a.w_gl = (const __nv_bfloat16*)w_hh_bf16;
    {
        const char* e = getenv("LAS_REC_WTMEM");
        // W_hh slice as the A operand from tensor memory: with the elected-lane issue + 4 accumulators the 32 UMMAs of a step
        // take 840 cycles instead of 1600 (shared-memory A: 4 KB of operand reads per UMMA bound it); LAS_REC_WTMEM=0 disables
        a.w_tmem = (e ? atoi(e) != 0 : true) && H <= 512;
    }
    LAS_CHECK_ARG((size_t)ndir * p.nslices * sizeof(unsigned) <= 1024, "lstm_rec_fwd_tc: too many batch slices");
    CUtensorMap tmW, tmH;
    rc = make_map_2d(&tmW, w_hh_bf16, H, (long long)ndir * 4 * H, 64, 32);
    if (rc) return rc;
    {   // h exchange buffer viewed as (64 k, rows, H/64 k-blocks): one box = the whole 32 x H slice in k-block-major smem order
        const int KBt = H / 64, KBPt = KBt / rec_pieces(KBt);
        const long long dims[3] = {64, (long long)ndir * 2 * p.Bpad, KBt};
        const long long strides[2] = {H, 64};
        const int box[3] = {64, NB_SLICE, KBPt};      // one piece of the tile per TMA issue (see the kernel)
        rc = make_map_nd(&tmH, a.hbuf, 3, dims, strides, box);
        if (rc) return rc;
    }
    // ---- DSMEM exchange (default when one chain per CTA; LAS_REC_DSMEM=0 disables): one cluster per group, h_t pushed SM-to-SM.
    // Measured 3.37 us/step against 3.54-3.64 for the counter/TMA exchange at B=96, H=512: the tensor pipe gets its operand
    // ~1300 cycles earlier; the outgoing copies share the SM's memory pipeline with the epilogue's own loads / stores, which is
    // why the helper warps take over the x-gate staging and the gate saves.  Clusters are independent of one
    // another, so a launch that cannot place all of them at once is still correct; any launch failure falls back below. ----
    {
        const char* de = getenv("LAS_REC_DSMEM");
        const size_t dsmem = 1024 + (size_t)p.chains * 2 * p.rs * 2048 + (size_t)p.chains * 2048 + 3 * 4 * 32 * 32 * 4 + 8 * 13 + 64;
        // (two chains per CTA, i.e. more batch slices than clusters fit: measured far slower than the TMA kernel -- one chain only)
        // at most 6 clusters: with 8 (B = 128 bidirectional) every GPC exchanges at once and a step takes 4.9 us instead of 2.5
        const bool few = p.nslices * ndir <= 6 || (de && atoi(de) == 2);
        if ((!de || atoi(de) != 0) && few && p.chains == 1 && p.rs <= 16 && H <= 512 && dsmem <= (size_t)las_device_info()->max_smem_optin) {
            auto kd = lstm_rec_fwd_dsm_kernel;
            cudaError_t e1 = cudaFuncSetAttribute(kd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsmem);
            if (e1 == cudaSuccess && p.rs > 8) e1 = cudaFuncSetAttribute(kd, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
            if (e1 == cudaSuccess) {
                cudaLaunchConfig_t cfg{};
                cfg.gridDim = dim3(p.rs, p.bsg, ndir); cfg.blockDim = dim3(NTHREADS8); cfg.dynamicSmemBytes = dsmem; cfg.stream = st;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeClusterDimension;
                at[0].val.clusterDim.x = p.rs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                cfg.attrs = at; cfg.numAttrs = 1;
                int nclusters = 0;
                e1 = cudaOccupancyMaxActiveClusters(&nclusters, kd, &cfg);
                if (e1 == cudaSuccess && nclusters < 1) e1 = cudaErrorInvalidConfiguration;
                if (e1 == cudaSuccess) {
                    LasProfScope prof(LAS_PROF_REC_FWD, stream, (double)T);
                    const bool publish = prog_ctr && prog_every > 0 && p.bsg * ndir <= 64;
                    a.progress = publish ? prog_ctr : nullptr; a.progress_every = prog_every;
                    e1 = cudaLaunchKernelEx(&cfg, kd, a);
                    if (e1 == cudaSuccess) {
                        if (publish) { t_prog_clusters = p.bsg * ndir; t_prog_rs = p.rs; }
                        las_count_launch(1);
                        return LAS_OK;
                    }
                    a.progress = nullptr;
                }
            }
            cudaGetLastError();         // fall through to the global-memory exchange
        }
    }
    LAS_CUDA(cudaMemsetAsync(ws, 0, 1024, st));
    LasProfScope prof(LAS_PROF_REC_FWD, stream, (double)T);
    void* kern = a.w_tmem ? (void*)lstm_rec_fwd_tc_kernel<true> : (void*)lstm_rec_fwd_tc_kernel<false>;
    LAS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
    void* args[] = {(void*)&tmW, (void*)&tmH, (void*)&a};
    LAS_CUDA(cudaLaunchCooperativeKernel(kern, dim3(p.rs, p.bsg, ndir), dim3(NTHREADS8), args, p.smem, st));
    las_count_launch(1);
    return LAS_OK;
}

extern "C" int las_lstm_rec_fwd_tc(float* gates, const void* w_hh_bf16, const int* lens, const float* drop_mask, float* out,
                                   float* hs_pad, float* cs_pad, int B, int T, int H, int ndir, int save_gates, void* ws, size_t ws_bytes,
                                   void* stream) {
    LAS_CHECK_ARG(hs_pad, "lstm_rec_fwd_tc: null pointer");
    return rec_fwd_tc_impl(gates, w_hh_bf16, lens, drop_mask, out, hs_pad, cs_pad, B, T, H, ndir, save_gates, ws, ws_bytes, nullptr, nullptr, stream);
}

// Same, plus bf16 copies of the two outputs that only GEMMs read afterwards: `out_bf16` (B, T, ndir*H) = the layer output (after
// locked dropout) as the next layer's / key_map's / value_map's A operand, `hs_bf16` (B, T+2, ndir*H) = the zero-framed hidden
// states as the dW_hh operand of backward.  Either may be null; `hs_pad` (fp32) may be null when `hs_bf16` is given, `out` when
// nobody reads the fp32 output.  Saves one fp32 read + bf16 write pass over (B, T, ndir*H) per copy.
extern "C" int las_lstm_rec_fwd_tc_ex(float* gates, const void* w_hh_bf16, const int* lens, const float* drop_mask, float* out,
                                      float* hs_pad, float* cs_pad, int B, int T, int H, int ndir, int save_gates, void* ws,
                                      size_t ws_bytes, void* out_bf16, void* hs_bf16, void* stream) {
    return rec_fwd_tc_impl(gates, w_hh_bf16, lens, drop_mask, out, hs_pad, cs_pad, B, T, H, ndir, save_gates, ws, ws_bytes,
                           (__nv_bfloat16*)out_bf16, (__nv_bfloat16*)hs_bf16, stream);
}

// =====================================================================================================================
// BPTT on the tensor pipe.
//   dh_rec[b, u] = sum_r dG_{next}[b, r] * W_hh[r, u]   (r over all 4H gate rows)   -- the serial dependency of backward.
// CTA (r, s, dir) owns the same 32 units / batch slice as in forward.  Resident in shared memory: the 32 x 4H slice of
// W_hh^T (bf16, K-major) as the B operand (N = 32 units).  Per step the 32 x 4H bf16 slice of the previous step's
// d(pre-activation) arrives by TMA straight from the (B*T, ndir*4H) bf16 dG array -- which is also exactly what the
// dX / dW_ih / dW_hh tensor-core GEMMs consume afterwards -- in 4 chunks through a 2-stage ring, as the A operand.
// UMMA M = 64 with 32 real batch rows: descriptors of rows 32-63 run into the next k-block's bytes; those accumulator
// lanes are never read.  The pointwise LSTM backward (gate derivatives, dc carry in registers) is fused in the epilogue.
// =====================================================================================================================
namespace {

struct RecTcBwdArgs {
    float* gates;            // (B, T, ndir, 4H) in: activated gates ; out: d(pre-activation), fp32
    __nv_bfloat16* dgb;      // (B*T, ndir*4H) bf16 copy of the same (TMA source + GEMM operand)
    const float* dout;       // (B, T, ndir*H)
    const float* cs_pad;     // (B, T+2, ndir*H)
    const int* lens;
    const float* mask;
    unsigned* ctr;
    int B, T, H, ndir, nslices, chains, bsg, KBr, CH, Bpad;
    __nv_bfloat16* dgx;      // K-split variant: compact bf16 exchange buffer (ndir, 2, 4 gates, Bpad, H)
    const __nv_bfloat16* w_t; // W_hh^T bf16 (ndir, H, 4H): source of the TMEM-resident A operand
    float* dbp;               // optional (ndir, nslices, 4H): per-batch-slice bias-gradient partial sums; when given the fp32
                              // d(pre-activation) write-back into `gates` is skipped (its only reader was the bias column sum)
    long long* dbg;
    unsigned* start_ctr;      // optional: every CTA adds 1 when it starts (a second stream waits for the sum, see las_set_launch_start_stream)
    unsigned* progress;       // optional (DSMEM kernel): per-cluster progress words, see las_lstm_rec_bwd_arm_progress
    int progress_every;
};

constexpr uint32_t IDESC_BWD = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(UNITS >> 3) << 17) | ((uint32_t)(64 >> 4) << 24);
constexpr int BWD_STAGES = 2;

__global__ void __launch_bounds__(NTHREADS, 1) lstm_rec_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmWt,
                                                                     const __grid_constant__ CUtensorMap tmG, const RecTcBwdArgs a) {
    extern __shared__ uint8_t smem_raw[];
    const int H = a.H, T = a.T, KBr = a.KBr, CH = a.CH, NCHUNK = KBr / CH;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t b_sm = base;                                        // KBr x [32 rows(u) x 128 B]   resident W_hh^T slice
    const uint32_t a_sm = b_sm + KBr * 4096;                           // BWD_STAGES x CH x [32 rows(b) x 128 B] (+4 KB slack)
    const uint32_t ex_off = (a_sm - smem_u32(smem_raw)) + BWD_STAGES * CH * 4096 + 4096;
    float* exD = reinterpret_cast<float*>(smem_raw + ex_off);          // [32][33]
    const uint32_t bar_base = smem_u32(smem_raw) + ex_off + 32 * 33 * 4 + 8;
    const uint32_t bar_al = (bar_base + 7u) & ~7u;
    auto full_bar = [&](int s) { return bar_al + 8u * s; };
    auto empty_bar = [&](int s) { return bar_al + 8u * (BWD_STAGES + s); };
    auto tfull_bar = [&](int c) { return bar_al + 8u * (2 * BWD_STAGES + c); };
    const uint32_t wbar = bar_al + 8u * (2 * BWD_STAGES + MAX_CHAINS);
    const uint32_t tmem_slot = bar_al + 8u * (2 * BWD_STAGES + MAX_CHAINS + 1);
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int r = blockIdx.x, sg = blockIdx.y, dir = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int F = a.ndir * H, G4 = 4 * H, NG = a.ndir * G4;
    const long long brow = (long long)(T + 2) * F;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmWt) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmG) : "memory");
        for (int s = 0; s < BWD_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int c = 0; c < MAX_CHAINS; ++c) mbar_init(tfull_bar(c), 1);
        mbar_init(wbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(64) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(wbar, (uint32_t)KBr * 4096u);
            for (int kb = 0; kb < KBr; ++kb) tma_load_2d(b_sm + kb * 4096, &tmWt, wbar, kb * 64, dir * H + r * UNITS);
            int stage = 0; uint32_t phase = 0;
            for (int s = 1; s < T; ++s) {
                const int t_prev = (dir == 0) ? (T - s) : (s - 1);        // time index processed at step s-1
                for (int c = 0; c < a.chains; ++c) {
                    const int slice = sg + c * a.bsg;
                    if (slice >= a.nslices) continue;
                    const unsigned* ctr = a.ctr + dir * a.nslices + slice;
                    const unsigned target = (unsigned)(H / UNITS) * (unsigned)s;
                    while (ld_acquire_gpu(ctr) < target) { }
                    asm volatile("fence.proxy.async;" ::: "memory");
                    for (int ch = 0; ch < NCHUNK; ++ch) {
                        mbar_wait(empty_bar(stage), phase ^ 1u);
                        mbar_arrive_expect_tx(full_bar(stage), (uint32_t)CH * 4096u);
                        // box (64, 32 batch rows, CH k-blocks, 1 timestep): one issue per chunk
                        tma_load_4d(a_sm + stage * CH * 4096, &tmG, full_bar(stage), 0, slice * NB_SLICE, dir * KBr + ch * CH, t_prev);
                        if (++stage == BWD_STAGES) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            mbar_wait(wbar, 0);
            int stage = 0; uint32_t phase = 0;
            for (int s = 1; s < T; ++s) {
                for (int c = 0; c < a.chains; ++c) {
                    const int slice = sg + c * a.bsg;
                    if (slice >= a.nslices) continue;
                    const uint32_t d_tmem = tmem_base + c * UNITS;
                    for (int ch = 0; ch < NCHUNK; ++ch) {
                        mbar_wait(full_bar(stage), phase);
                        tc_fence_after();
                        for (int kl = 0; kl < CH; ++kl) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const uint64_t ad = make_desc_k(a_sm + (stage * CH + kl) * 4096 + k * 32);
                                const uint64_t bd = make_desc_k(b_sm + (ch * CH + kl) * 4096 + k * 32);
                                umma_bf16(d_tmem, ad, bd, IDESC_BWD, (ch | kl | k) ? 1u : 0u);
                            }
                        }
                        umma_commit(empty_bar(stage));
                        if (++stage == BWD_STAGES) { stage = 0; phase ^= 1u; }
                    }
                    umma_commit(tfull_bar(c));
                }
            }
        }
    } else if (warp >= 4) {
        const int q = warp & 3, j = lane;
        const int te = (warp - 4) * 32 + lane;
        const int u = r * UNITS + j;
        float dcst[MAX_CHAINS][8];
        int lenr[MAX_CHAINS][8];
#pragma unroll
        for (int c = 0; c < MAX_CHAINS; ++c)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                dcst[c][i] = 0.f;
                const int b = (sg + c * a.bsg) * NB_SLICE + q * 8 + i;
                lenr[c][i] = (c < a.chains && sg + c * a.bsg < a.nslices && b < a.B) ? a.lens[b] : 0;
            }
        for (int s = 0; s < T; ++s) {
            const int t = (dir == 0) ? (T - 1 - s) : s;
            const int fprev = (dir == 0) ? t : t + 2, fcur = t + 1;
#pragma unroll
            for (int c = 0; c < MAX_CHAINS; ++c) {
                const int slice = sg + c * a.bsg;
                if (c >= a.chains || slice >= a.nslices) continue;
                const int b0 = slice * NB_SLICE;
                // operands of the pointwise backward, issued before waiting on the tensor pipe
                float gi[8], gf[8], gg[8], go[8], ct[8], cp[8], dh[8], mk[8], rec[8];
                bool valid[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int b = b0 + q * 8 + i;
                    valid[i] = t < lenr[c][i];
                    // pure loads, no branches and no arithmetic: an in-order warp would otherwise stall on the first use and
                    // serialise the eight rows' HBM latencies.  Rows past B are clamped to a valid address; invalid rows are
                    // zeroed in the pointwise step.
                    const int bc = b < a.B ? b : a.B - 1;
                    const float* gp = a.gates + (((long long)bc * T + t) * a.ndir + dir) * G4 + u;
                    gi[i] = gp[0]; gf[i] = gp[H]; gg[i] = gp[2 * H]; go[i] = gp[3 * H];
                    ct[i] = a.cs_pad[(long long)bc * brow + (long long)fcur * F + dir * H + u];
                    cp[i] = a.cs_pad[(long long)bc * brow + (long long)fprev * F + dir * H + u];
                    mk[i] = a.mask ? a.mask[(long long)bc * F + dir * H + u] : 1.f;
                    dh[i] = a.dout[((long long)bc * T + t) * F + dir * H + u];
                    rec[i] = 0.f;
                }
                if (s > 0) {
                    mbar_wait(tfull_bar(c), (uint32_t)((s - 1) & 1));
                    tc_fence_after();
                    if (q < 2) {
                        // UMMA M = 64: accumulator rows 0-15 sit in TMEM lanes 0-15, rows 16-31 in lanes 32-47
                        uint32_t v[32];
                        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + c * UNITS, v);
                        if (lane < 16) {
#pragma unroll
                            for (int n = 0; n < 32; ++n) exD[(q * 16 + lane) * 33 + n] = __uint_as_float(v[n]);
                        }
                    }
                    tc_fence_before();
                    named_bar_sync(1, 128);
#pragma unroll
                    for (int i = 0; i < 8; ++i) rec[i] = exD[(q * 8 + i) * 33 + j];
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int b = b0 + q * 8 + i;
                    if (b >= a.B) continue;
                    float dai = 0.f, daf = 0.f, dag = 0.f, dao = 0.f, dcn = 0.f;
                    if (valid[i]) {
                        const float tcv = tanh_fast(ct[i]);
                        const float dhv = fmaf(dh[i], mk[i], rec[i]);
                        const float dct = fmaf(dhv * go[i], 1.f - tcv * tcv, dcst[c][i]);
                        dai = dct * gg[i] * gi[i] * (1.f - gi[i]);
                        daf = dct * cp[i] * gf[i] * (1.f - gf[i]);
                        dag = dct * gi[i] * (1.f - gg[i] * gg[i]);
                        dao = dhv * tcv * go[i] * (1.f - go[i]);
                        dcn = dct * gf[i];
                    }
                    dcst[c][i] = dcn;
                    const long long row = (long long)b * T + t;
                    // peers (and the GEMMs afterwards) read the bf16 copy; it is what the publish below releases
                    __nv_bfloat16* bp = a.dgb + row * NG + dir * G4 + u;
                    bp[0] = __float2bfloat16(dai); bp[H] = __float2bfloat16(daf);
                    bp[2 * H] = __float2bfloat16(dag); bp[3 * H] = __float2bfloat16(dao);
                    gi[i] = dai; gf[i] = daf; gg[i] = dag; go[i] = dao;
                }
                named_bar_sync(1, 128);
                if (te == 0 && s + 1 < T) red_release_gpu_add(a.ctr + dir * a.nslices + slice, 1u);
                // fp32 d(pre-activation) (bias-gradient column sums read it) after the publish
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int b = b0 + q * 8 + i;
                    if (b >= a.B) continue;
                    float* gp = a.gates + ((((long long)b * T + t) * a.ndir + dir) * G4) + u;
                    gp[0] = gi[i]; gp[H] = gf[i]; gp[2 * H] = gg[i]; gp[3 * H] = go[i];
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64) : "memory");
    }
}


// dst[b][c][r] (bf16) = src[b][r][c] (fp32): W_hh (ndir, 4H, H) -> W_hh^T (ndir, H, 4H)
__global__ void __launch_bounds__(256) transpose_cast_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int rows, int cols) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z, r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const float* s = src + (long long)b * rows * cols;
    __nv_bfloat16* d = dst + (long long)b * rows * cols;
    for (int i = ty; i < 32; i += 8) tile[i][tx] = (r0 + i < rows && c0 + tx < cols) ? s[(long long)(r0 + i) * cols + c0 + tx] : 0.f;
    __syncthreads();
    for (int i = ty; i < 32; i += 8)
        if (c0 + i < cols && r0 + tx < rows) d[(long long)(c0 + i) * rows + r0 + tx] = __float2bfloat16(tile[tx][i]);
}

}  // namespace

extern "C" int las_transpose_cast_bf16(const float* src, void* dst, int batch, int rows, int cols, void* stream) {
    LAS_CHECK_ARG(src && dst && batch >= 1 && rows >= 1 && cols >= 1, "transpose_cast: bad arguments");
    int rc = las_set_device_of(dst);
    if (rc) return rc;
    transpose_cast_kernel<<<dim3(ceil_div(cols, 32), ceil_div(rows, 32), batch), 256, 0, (cudaStream_t)stream>>>(src, (__nv_bfloat16*)dst, rows, cols);
    LAS_LAUNCH_CHECK();
    return LAS_OK;
}

// ---- "start" hand-off to a second stream (las_set_launch_start_stream) ----
// The next backward launch of this thread releases `t_start_stream` -- a stream other than the one the kernel runs on -- as soon
// as every CTA of the recurrence kernel is resident: each CTA bumps a device counter on entry and the side stream sits in a
// cuStreamWaitValue32 on it.  Work queued on that stream afterwards (the weight-gradient GEMMs of the layer above, capped to the
// SMs the recurrence leaves free) then runs BESIDE the latency-bound recurrence without having delayed its cluster placement.
// (A programmatic launch event with triggerAtBlockStart was tried first: a stream waiting on it was only released when the
// kernel had finished.)  On the variants that do not carry the counter, or if the driver refuses the wait, the side stream
// waits for the kernel's completion instead -- correct, without overlap; las_launch_start_mode() tells which.
static thread_local cudaStream_t t_start_stream = nullptr;
static thread_local bool t_start_armed = false;
static thread_local int t_start_mode = 0;
extern "C" void las_set_launch_start_stream(void* side_stream) {
    t_start_stream = (cudaStream_t)side_stream;
    t_start_armed = side_stream != nullptr;
    t_start_mode = 0;
}
extern "C" int las_launch_start_mode(void) { return t_start_mode; }

namespace {
std::mutex g_start_mu;
unsigned* g_start_ctr[64];          // per device: CTA-start counter (never reset: targets advance, the comparison is cyclic)
unsigned g_start_target[64];
}  // namespace

// an armed hand-off that no launcher consumed: release the side stream when everything enqueued on `st` so far has finished
struct StartEventScope {
    cudaStream_t st;
    bool outer;
    explicit StartEventScope(cudaStream_t s, bool o) : st(s), outer(o) {}
    ~StartEventScope() {
        if (!outer || !t_start_armed) return;
        t_start_armed = false;
        static thread_local cudaEvent_t ev = nullptr;
        if (!ev && cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); ev = nullptr; }
        if (ev && cudaEventRecord(ev, st) == cudaSuccess && cudaStreamWaitEvent(t_start_stream, ev, 0) == cudaSuccess) return;
        cudaGetLastError();
        cudaStreamSynchronize(st);          // last resort: the host waits
    }
};

static int launch_bwd_tc2(const float* dout, float* gates, void* dgates_bf16, const float* cs_pad, const void* w_hh_t_bf16, const int* lens,
                          const float* drop_mask, int B, int T, int H, int ndir, void* ws, cudaStream_t st, void* stream, float* dbp);
static int launch_bwd_dsm(const float* dout, float* gates, void* dgates_bf16, const float* cs_pad, const void* w_hh_t_bf16, const int* lens,
                          const float* drop_mask, int B, int T, int H, int ndir, cudaStream_t st, void* stream, float* dbp);

static int rec_bwd_tc_impl(const float* dout, float* gates, void* dgates_bf16, const float* cs_pad, const void* w_hh_t_bf16,
                           const int* lens, const float* drop_mask, int B, int T, int H, int ndir, void* ws, size_t ws_bytes,
                           void* stream, float* dbp);

extern "C" int las_lstm_rec_bwd_tc(const float* dout, float* gates, void* dgates_bf16, const float* cs_pad, const void* w_hh_t_bf16,
                                   const int* lens, const float* drop_mask, int B, int T, int H, int ndir, void* ws, size_t ws_bytes,
                                   void* stream) {
    return rec_bwd_tc_impl(dout, gates, dgates_bf16, cs_pad, w_hh_t_bf16, lens, drop_mask, B, T, H, ndir, ws, ws_bytes, stream, nullptr);
}

extern "C" int las_lstm_rec_bwd_tc_dbias_slices(int B, int H, int ndir) {
    // > 0: the K-split BPTT kernel runs for this shape and can emit per-batch-slice bias-gradient rows; 0: not available
    if (H % 128 != 0 || B < 1 || (ndir != 1 && ndir != 2)) return 0;
    const char* e = getenv("LAS_REC_BWD_KSPLIT");
    if (e && atoi(e) == 0) return 0;
    const int rs = 4 * (H / 128), nslices = ceil_div(B, NB_SLICE);
    const int max_bsg = las_device_info()->num_sms / (rs * ndir);
    if (max_bsg < 1) return 0;
    const int bsg = nslices < max_bsg ? nslices : max_bsg;
    if (ceil_div(nslices, bsg) > MAX_CHAINS) return 0;
    return nslices;
}

extern "C" int las_lstm_rec_bwd_tc_db(const float* dout, float* gates, void* dgates_bf16, const float* cs_pad, const void* w_hh_t_bf16,
                                      const int* lens, const float* drop_mask, int B, int T, int H, int ndir, void* ws, size_t ws_bytes,
                                      float* dbias_partial, void* stream) {
    LAS_CHECK_ARG(dbias_partial != nullptr, "lstm_rec_bwd_tc_db: null dbias_partial");
    LAS_CHECK_ARG(las_lstm_rec_bwd_tc_dbias_slices(B, H, ndir) > 0, "lstm_rec_bwd_tc_db: not available for B=%d H=%d (use las_lstm_rec_bwd_tc)", B, H);
    return rec_bwd_tc_impl(dout, gates, dgates_bf16, cs_pad, w_hh_t_bf16, lens, drop_mask, B, T, H, ndir, ws, ws_bytes, stream, dbias_partial);
}

static int rec_bwd_tc_impl(const float* dout, float* gates, void* dgates_bf16, const float* cs_pad, const void* w_hh_t_bf16,
                           const int* lens, const float* drop_mask, int B, int T, int H, int ndir, void* ws, size_t ws_bytes,
                           void* stream, float* dbp) {
    LAS_CHECK_ARG(dout && gates && dgates_bf16 && cs_pad && w_hh_t_bf16 && lens && ws, "lstm_rec_bwd_tc: null pointer");
    LAS_CHECK_ARG(B >= 1 && T >= 1 && (ndir == 1 || ndir == 2), "lstm_rec_bwd_tc: bad dims");
    int rc = las_set_device_of(gates);
    if (rc) return rc;
    Plan p;
    rc = make_plan(B, H, ndir, &p);
    if (rc) return rc;
    if (ws_bytes < las_lstm_rec_tc_workspace_bytes(B, H, ndir)) { las_set_error("lstm_rec_bwd_tc: workspace too small"); return LAS_ERR_WORKSPACE; }
    cudaStream_t st = (cudaStream_t)stream;
    struct DisarmProgress { ~DisarmProgress() { t_bprog_ctr = nullptr; t_bprog_every = 0; } } disarm_progress;      // the request is for this call only
    static thread_local int depth = 0;
    struct Depth { Depth() { ++depth; } ~Depth() { --depth; } } depth_scope;
    StartEventScope start_event_scope(st, depth == 1);
    {
        // same batch passes as the forward launcher when the rows do not fit one chain per CTA
        const char* sb = getenv("LAS_REC_SPLIT_BATCH");
        const char* de = getenv("LAS_REC_DSMEM");
        if (p.chains > 1 && !dbp && (sb && atoi(sb) == 1) && !(de && atoi(de) == 0) && H % 128 == 0 && H <= 512) {
            t_bprog_ctr = nullptr;          // per-pass launches: no whole-batch progress
            const int rows = p.bsg * NB_SLICE;
            const long long F = (long long)ndir * H;
            for (int b0 = 0; b0 < B; b0 += rows) {
                const int nb = B - b0 < rows ? B - b0 : rows;
                rc = rec_bwd_tc_impl(dout + (long long)b0 * T * F, gates + (long long)b0 * T * ndir * 4 * H,
                                     (__nv_bfloat16*)dgates_bf16 + (long long)b0 * T * ndir * 4 * H, cs_pad + (long long)b0 * (T + 2) * F, w_hh_t_bf16,
                                     lens + b0, drop_mask ? drop_mask + b0 * F : nullptr, nb, T, H, ndir, ws, ws_bytes, stream, nullptr);
                if (rc) return rc;
            }
            return LAS_OK;
        }
    }
    {
        const char* e = getenv("LAS_REC_BWD_KSPLIT");
        if (!e || atoi(e) != 0) {
            rc = launch_bwd_dsm(dout, gates, dgates_bf16, cs_pad, w_hh_t_bf16, lens, drop_mask, B, T, H, ndir, st, stream, dbp);
            if (rc == LAS_OK) return LAS_OK;
            rc = launch_bwd_tc2(dout, gates, dgates_bf16, cs_pad, w_hh_t_bf16, lens, drop_mask, B, T, H, ndir, ws, st, stream, dbp);
            if (rc == LAS_OK) return LAS_OK;          // otherwise fall through to the streaming variant
            if (dbp) return rc;                       // (the streaming variant has no bias-gradient output)
        }
    }
    RecTcBwdArgs a{};
    a.gates = gates; a.dgb = (__nv_bfloat16*)dgates_bf16; a.dout = dout; a.cs_pad = cs_pad; a.lens = lens; a.mask = drop_mask;
    a.ctr = (unsigned*)ws;
    a.B = B; a.T = T; a.H = H; a.ndir = ndir; a.nslices = p.nslices; a.chains = p.chains; a.bsg = p.bsg;
    a.KBr = 4 * H / 64; a.CH = a.KBr < 8 ? a.KBr : 8;
    const size_t smem = 1024 + (size_t)a.KBr * 4096 + (size_t)BWD_STAGES * a.CH * 4096 + 4096 + 32 * 33 * 4 + 16 +
                        8 * (2 * BWD_STAGES + MAX_CHAINS + 2) + 64;
    LAS_CHECK_ARG(smem <= (size_t)las_device_info()->max_smem_optin, "lstm_rec_bwd_tc: needs %zu B of shared memory", smem);
    const long long NG = (long long)ndir * 4 * H;
    CUtensorMap tmWt, tmG;
    rc = make_map_2d(&tmWt, w_hh_t_bf16, 4LL * H, (long long)ndir * H, 64, 32);
    if (rc) return rc;
    {   // dG (B*T, NG) viewed as (64 r, B rows [stride T*NG], NG/64 k-blocks [stride 64], T [stride NG])
        const long long dims[4] = {64, B, NG / 64, T};
        const long long strides[3] = {(long long)T * NG, 64, NG};
        const int box[4] = {64, NB_SLICE, a.CH, 1};
        rc = make_map_nd(&tmG, dgates_bf16, 4, dims, strides, box);
        if (rc) return rc;
    }
    LAS_CUDA(cudaFuncSetAttribute(lstm_rec_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LAS_CUDA(cudaMemsetAsync(ws, 0, 1024, st));
    LasProfScope prof(LAS_PROF_REC_BWD, stream, (double)T);
    void* args[] = {(void*)&tmWt, (void*)&tmG, (void*)&a};
    LAS_CUDA(cudaLaunchCooperativeKernel((void*)lstm_rec_bwd_tc_kernel, dim3(p.rs, p.bsg, ndir), dim3(NTHREADS), args, smem, st));
    las_count_launch(1);
    return LAS_OK;
}

// =====================================================================================================================
// BPTT, K-split variant (H % 128 == 0): the default.
// The per-step product dh_rec = dG . W_hh has a 4H-long reduction, 4x the forward's.  Instead of streaming the whole
// 32 x 4H dG slice through every CTA, the reduction is split by GATE over a 4-CTA thread-block cluster:
//   CTA (kq, ub, s, dir), kq = cluster rank = gate index, ub = 128-unit block.
//   resident A operand : W_hh^T[units of block ub, rows of gate kq]           128 x H bf16  (K-major, shared memory)
//   per-step B operand : dG_prev[32 batch rows, gate kq, all H units]          32 x H bf16  (one TMA issue, same size as fwd)
//   UMMA M = 128, N = 32, K = H  ->  partial dh_rec for 128 units x 32 rows over one gate, in TMEM.
// The four partials are summed through distributed shared memory (each CTA finalises 32 of the block's 128 units), then
// the pointwise LSTM backward runs for those 32 units exactly as in the other variant.  Cross-cluster ordering (a CTA's B
// operand spans all units, i.e. all 4*H/128 CTAs of the (direction, batch slice) group) still uses the release/acquire counter.
// =====================================================================================================================
namespace {

__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ float ld_dsmem_f32(uint32_t local_saddr, uint32_t cta) {
    uint32_t raddr;
    float v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(local_saddr), "r"(cta));
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(raddr));
    return v;
}

constexpr int PART_LD = 132;          // floats per batch row of the partial tile [32 b][128 u] (+4 pad)

constexpr int BWD_NACC = 4;

template <bool WTMEM>
__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(NTHREADS, 1)
    lstm_rec_bwd_tc2_kernel(const __grid_constant__ CUtensorMap tmWt, const __grid_constant__ CUtensorMap tmG, const RecTcBwdArgs a) {
    extern __shared__ uint8_t smem_raw[];
    const int H = a.H, T = a.T, KB = H / 64;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_sm = base;                                        // KB x [128 rows(u) x 128 B]  resident W_hh^T tile
    const uint32_t b_sm = a_sm + (WTMEM ? 0 : KB * 16384);             // chains x KB x [32 rows(b) x 128 B]
    const uint32_t part_off = (b_sm - smem_u32(smem_raw)) + a.chains * KB * 4096;
    // partial-sum exchange inside the 4-CTA cluster, pushed (no cluster barrier, no remote loads): stg[parity][piece q] = this
    // CTA's partial for the 32 units CTA q finalises ([32 rows][32 units] fp32, 4 KB), rcv[parity][src] = what the four CTAs sent me
    float* part0 = reinterpret_cast<float*>(smem_raw + part_off);      // stg: [2][4][32][32]
    float* rcv0 = part0 + 2 * 4 * 1024;                                // rcv: [2][4][32][32]
    const uint32_t part_saddr0 = smem_u32(smem_raw) + part_off;
    const uint32_t rcv_saddr0 = part_saddr0 + 2 * 4 * 4096;
    const uint32_t bar_base = (part_saddr0 + 4 * 4 * 4096 + 15u) & ~15u;
    auto full_bar = [&](int c) { return bar_base + 8u * c; };
    auto tfull_bar = [&](int c) { return bar_base + 8u * (MAX_CHAINS + c); };
    const uint32_t wbar = bar_base + 8u * (2 * MAX_CHAINS);
    const uint32_t tmem_slot = bar_base + 8u * (2 * MAX_CHAINS + 1);
    const int NP = rec_pieces(KB), KBP = KB / NP;                 // the dG tile arrives in pieces (see the forward kernel)
    auto piece_bar = [&](int c, int pc) { return pc == 0 ? full_bar(c) : bar_base + 8u * (2 * MAX_CHAINS + 2 + (pc - 1) * MAX_CHAINS + c); };
    auto rbar = [&](int par) { return bar_base + 8u * (2 * MAX_CHAINS + 2 + 3 * MAX_CHAINS + par); };      // partials received
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int kq = (int)cluster_ctarank();            // gate handled by this CTA's reduction slice
    const int ub = blockIdx.x >> 2;                   // 128-unit block
    const int sg = blockIdx.y, dir = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int F = a.ndir * H, G4 = 4 * H, NG = a.ndir * G4;
    const long long brow = (long long)(T + 2) * F;
    const unsigned group = (unsigned)(4 * (H / 128));  // CTAs sharing one (direction, batch slice)

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmWt) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmG) : "memory");
        for (int c = 0; c < MAX_CHAINS; ++c) {
            mbar_init(tfull_bar(c), 1);
            for (int pc = 0; pc < 4; ++pc) mbar_init(piece_bar(c, pc), 1);
        }
        mbar_init(wbar, 1);
        mbar_init(rbar(0), 1); mbar_init(rbar(1), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // TMEM map: BWD_NACC independent accumulators per chain first, then (WTMEM) the resident W_hh^T tile (H/2 columns)
    const uint32_t tmem_cols = WTMEM ? 512u : 256u;
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    const uint32_t tmem_w = tmem_base + MAX_CHAINS * BWD_NACC * NB_SLICE;
    if (WTMEM) {
        if (warp >= 4) {
            // A operand in tensor memory: TMEM lane = unit row of the block, column pair = two consecutive k (gate rows of gate kq)
            const int qq = warp & 3;
            const uint32_t* wrow = reinterpret_cast<const uint32_t*>(a.w_t + ((long long)(dir * H + ub * 128 + qq * 32 + lane)) * G4 + kq * H);
            for (int cb = 0; cb < H / 64; ++cb) {
                uint32_t v[32];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const uint4 t4 = *reinterpret_cast<const uint4*>(wrow + cb * 32 + i * 4);
                    v[i * 4 + 0] = t4.x; v[i * 4 + 1] = t4.y; v[i * 4 + 2] = t4.z; v[i * 4 + 3] = t4.w;
                }
                tmem_st32(tmem_w + ((uint32_t)(qq * 32) << 16) + cb * 32, v);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }

    cluster_sync_all();              // every CTA of the cluster has initialised its barriers before a peer pushes partials at it

    // per-thread state of the epilogue role (declared for all so the step loop below is shared by every warp)
    // Epilogue ownership: thread (warp q, lane) finalises FOUR consecutive units u0..u0+3 (u0 = block base + 4*(lane & 7)) for TWO
    // batch rows (q*8 + 2*(lane >> 3) + {0, 1}), so that every global access of the pointwise backward is a 128-bit one: a quarter
    // of the memory instructions of a one-unit-per-lane mapping (their issue rate, not the exchange, had become the critical path).
    const int q = warp & 3, j = lane;
    const int te = (warp - 4) * 32 + lane;
    const int jj = lane & 7, rr = lane >> 3;
    const int u0 = ub * 128 + kq * 32 + 4 * jj;        // first of the 4 units this thread finalises
    const int rl0 = q * 8 + rr * 2;                    // first of its 2 batch rows inside the 32-row slice
    float4 dcst[MAX_CHAINS][2];
    float4 mkr[MAX_CHAINS][2];                         // locked-dropout mask: constant over time
    float4 cnext[MAX_CHAINS][2];                       // c_{t-1} loaded at this step = c_t of the next step (time runs backwards)
    float4 dbacc[MAX_CHAINS][4];                       // bias-gradient partial sums (gate x 4 units) over this thread's rows, all steps
    int lenr[MAX_CHAINS][2];
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int c = 0; c < MAX_CHAINS; ++c) {
#pragma unroll
        for (int gq = 0; gq < 4; ++gq) dbacc[c][gq] = z4;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            dcst[c][i] = z4; cnext[c][i] = z4;
            const int b = (sg + c * a.bsg) * NB_SLICE + rl0 + i;
            lenr[c][i] = (warp >= 4 && c < a.chains && sg + c * a.bsg < a.nslices && b < a.B) ? a.lens[b] : 0;
            mkr[c][i] = (a.mask && lenr[c][i] > 0) ? *reinterpret_cast<const float4*>(a.mask + (long long)b * F + dir * H + u0)
                                                   : make_float4(1.f, 1.f, 1.f, 1.f);
        }
    }

    if (!WTMEM) {
        if (warp == 0 && lane == 0) {
            mbar_arrive_expect_tx(wbar, (uint32_t)KB * 16384u);
            for (int kb = 0; kb < KB; ++kb) tma_load_2d(a_sm + kb * 16384, &tmWt, wbar, kq * H + kb * 64, dir * H + ub * 128);
        }
        if (warp == 1) mbar_wait(wbar, 0);
    }

    // Every warp walks the same (step, chain) sequence: the cluster barrier of each iteration needs all threads of all 4 CTAs.
    int iter = 0, riter = 0;
    for (int s = 0; s < T; ++s) {
        const int t = (dir == 0) ? (T - 1 - s) : s;
        const int t_prev = (dir == 0) ? (T - s) : (s - 1);
        const int fprev = (dir == 0) ? t : t + 2, fcur = t + 1;
#pragma unroll
        for (int c = 0; c < MAX_CHAINS; ++c) {
            const int slice = sg + c * a.bsg;
            if (c >= a.chains || slice >= a.nslices) continue;          // uniform across the cluster (same sg, chains)
            const int b0 = slice * NB_SLICE;
            const int rpar = riter & 1;                                   // parity of this reduce-iteration (s > 0 only)
            const uint32_t rphase = (uint32_t)((riter >> 1) & 1);
            if (s > 0) ++riter;
            ++iter;
            float4 gi[2], gf[2], gg[2], go[2], ct[2], cp[2], dh[2], rec[2];
            bool valid[2];
            if (warp == 0) {
                if (lane == 0 && s > 0) {
                    const unsigned* ctr = a.ctr + dir * a.nslices + slice;
                    const unsigned target = group * (unsigned)s;
                    while (ld_acquire_gpu(ctr) < target) { }
                    REC_STAMP(0);
                    asm volatile("fence.proxy.async.global;" ::: "memory");
                    // compact exchange buffer (dir, parity, gate, Bpad, H): the 32 x H tile of gate kq is 32 contiguous rows
                    const int row0 = ((dir * 2 + ((s - 1) & 1)) * 4 + kq) * a.Bpad + b0;
                    for (int pc = 0; pc < NP; ++pc) {
                        mbar_arrive_expect_tx(piece_bar(c, pc), (uint32_t)KBP * 4096u);
                        tma_load_3d(b_sm + (c * KB + pc * KBP) * 4096, &tmG, piece_bar(c, pc), 0, row0, pc * KBP);
                    }
                    REC_STAMP(1);
                }
                __syncwarp();
            } else if (warp == 1) {
                if (s > 0) {
                    // whole warp waits, one elected lane issues (uniform control flow keeps the descriptors in uniform registers);
                    // BWD_NACC independent accumulators (one per k sub-step) break the accumulate-dependency chain
                    for (int pc = 0; pc < NP; ++pc) {
                        mbar_wait(piece_bar(c, pc), (uint32_t)((s - 1) & 1));
                        if (lane == 0 && pc == 0) REC_STAMP(2);
                        tc_fence_after();
                        if (elect_one()) {
                            const int kb_lo = pc * KBP, kb_hi = kb_lo + KBP;
                            for (int kb = kb_lo; kb < kb_hi; ++kb) {
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    const uint32_t d_tmem = tmem_base + (uint32_t)((c * BWD_NACC + k) * NB_SLICE);
                                    const uint64_t bd = make_desc_k(b_sm + (c * KB + kb) * 4096 + k * 32);
                                    if (WTMEM) {
                                        umma_bf16_ts(d_tmem, tmem_w + kb * 32 + k * 8, bd, IDESC, kb ? 1u : 0u);
                                    } else {
                                        const uint64_t ad = make_desc_k(a_sm + kb * 16384 + k * 32);
                                        umma_bf16(d_tmem, ad, bd, IDESC, kb ? 1u : 0u);
                                    }
                                }
                            }
                            if (kb_hi == KB) umma_commit(tfull_bar(c));
                        }
                        __syncwarp();
                    }
                    __syncwarp();
                    if (lane == 0) REC_STAMP(3);
                }
                __syncwarp();
            } else if (warp >= 4) {
                // operands of the pointwise backward for this thread's unit, issued before waiting on the tensor pipe
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int b = b0 + rl0 + i;
                    valid[i] = t < lenr[c][i];
                    // pure 128-bit loads, no branches: rows past B are clamped to a valid address; invalid rows are zeroed in the
                    // pointwise step
                    const int bc = b < a.B ? b : a.B - 1;
                    const float* gp = a.gates + (((long long)bc * T + t) * a.ndir + dir) * G4 + u0;
                    gi[i] = *reinterpret_cast<const float4*>(gp); gf[i] = *reinterpret_cast<const float4*>(gp + H);
                    gg[i] = *reinterpret_cast<const float4*>(gp + 2 * H); go[i] = *reinterpret_cast<const float4*>(gp + 3 * H);
                    // c_t: carried over from the previous step's c_{t-1} load (first step: loaded)
                    if (s == 0) ct[i] = *reinterpret_cast<const float4*>(a.cs_pad + (long long)bc * brow + (long long)fcur * F + dir * H + u0);
                    else ct[i] = cnext[c][i];
                    cp[i] = *reinterpret_cast<const float4*>(a.cs_pad + (long long)bc * brow + (long long)fprev * F + dir * H + u0);
                    dh[i] = *reinterpret_cast<const float4*>(a.dout + ((long long)bc * T + t) * F + dir * H + u0);
                    rec[i] = z4;
                }
                if (te == 0) REC_STAMP(4);
                if (s > 0) {
                    mbar_wait(tfull_bar(c), (uint32_t)((s - 1) & 1));
                    if (te == 0) REC_STAMP(5);
                    tc_fence_after();
                    float accv[32];
#pragma unroll
                    for (int acc = 0; acc < BWD_NACC; ++acc) {
                        uint32_t v[32];
                        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((c * BWD_NACC + acc) * NB_SLICE), v);
#pragma unroll
                        for (int n = 0; n < 32; ++n) accv[n] = acc ? accv[n] + __uint_as_float(v[n]) : __uint_as_float(v[n]);
                    }
                    // TMEM lane = unit (32q + lane) of the block, column = batch row.  Warp q holds exactly the 32 units CTA q of the
                    // cluster finalises: stage them as [row][unit] and push the 4 KB piece into CTA q's receive buffer
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the piece staged two iterations ago has been read
                    __syncwarp();
                    float* stg = part0 + (rpar * 4 + q) * 1024;
#pragma unroll
                    for (int n = 0; n < 32; ++n) stg[n * 32 + lane] = accv[n];
                    tc_fence_before();
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) {
                        const uint32_t dst = mapa_u32(rcv_saddr0 + (uint32_t)((rpar * 4 + kq) * 4096), (uint32_t)q);
                        bulk_copy_to_peer(dst, part_saddr0 + (uint32_t)((rpar * 4 + q) * 4096), 4096u, mapa_u32(rbar(rpar), (uint32_t)q));
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                    if (te == 0) {
                        mbar_arrive_expect_tx(rbar(rpar), 4u * 4096u);         // the four pieces for my 32 units
                        REC_STAMP(6);
                    }
                }
            }
            if (s > 0) {
                if (warp >= 4) {
                    mbar_wait(rbar(rpar), rphase);
                    if (te == 0) REC_STAMP(7);
                    const float* rcv = rcv0 + rpar * 4 * 1024;
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const int o = (rl0 + i) * 32 + 4 * jj;
                        const float4 p0 = *reinterpret_cast<const float4*>(rcv + o), p1 = *reinterpret_cast<const float4*>(rcv + 1024 + o);
                        const float4 p2 = *reinterpret_cast<const float4*>(rcv + 2048 + o), p3 = *reinterpret_cast<const float4*>(rcv + 3072 + o);
                        rec[i] = make_float4((p0.x + p1.x) + (p2.x + p3.x), (p0.y + p1.y) + (p2.y + p3.y), (p0.z + p1.z) + (p2.z + p3.z),
                                             (p0.w + p1.w) + (p2.w + p3.w));
                    }
                }
                if (warp == 4 && lane == 0) REC_STAMP(11);
            }
            if (warp >= 4) {
                auto pw = [&](float gi_, float gf_, float gg_, float go_, float ct_, float cp_, float dh_, float mk_, float rec_, float dc_,
                              bool ok, float& dai, float& daf, float& dag, float& dao, float& dcn) {
                    dai = daf = dag = dao = dcn = 0.f;
                    if (ok) {
                        const float tcv = tanh_fast(ct_);
                        const float dhv = fmaf(dh_, mk_, rec_);
                        const float dct = fmaf(dhv * go_, 1.f - tcv * tcv, dc_);
                        dai = dct * gg_ * gi_ * (1.f - gi_);
                        daf = dct * cp_ * gf_ * (1.f - gf_);
                        dag = dct * gi_ * (1.f - gg_ * gg_);
                        dao = dhv * tcv * go_ * (1.f - go_);
                        dcn = dct * gf_;
                    }
                };
                auto pack4 = [](const float4& v) {
                    const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
                    uint2 r;
                    r.x = *reinterpret_cast<const uint32_t*>(&lo); r.y = *reinterpret_cast<const uint32_t*>(&hi);
                    return r;
                };
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int b = b0 + rl0 + i;
                    float4 dI, dF, dG, dO, dC;
                    pw(gi[i].x, gf[i].x, gg[i].x, go[i].x, ct[i].x, cp[i].x, dh[i].x, mkr[c][i].x, rec[i].x, dcst[c][i].x, valid[i], dI.x, dF.x, dG.x, dO.x, dC.x);
                    pw(gi[i].y, gf[i].y, gg[i].y, go[i].y, ct[i].y, cp[i].y, dh[i].y, mkr[c][i].y, rec[i].y, dcst[c][i].y, valid[i], dI.y, dF.y, dG.y, dO.y, dC.y);
                    pw(gi[i].z, gf[i].z, gg[i].z, go[i].z, ct[i].z, cp[i].z, dh[i].z, mkr[c][i].z, rec[i].z, dcst[c][i].z, valid[i], dI.z, dF.z, dG.z, dO.z, dC.z);
                    pw(gi[i].w, gf[i].w, gg[i].w, go[i].w, ct[i].w, cp[i].w, dh[i].w, mkr[c][i].w, rec[i].w, dcst[c][i].w, valid[i], dI.w, dF.w, dG.w, dO.w, dC.w);
                    dcst[c][i] = dC;
                    cnext[c][i] = cp[i];
                    // what the peers' next step reads: bf16 d(pre-activation) in the compact exchange buffer, 4 units per 64-bit store
                    __nv_bfloat16* xp = a.dgx + ((long long)((dir * 2 + (s & 1)) * 4) * a.Bpad + b) * H + u0;
                    const long long gst = (long long)a.Bpad * H;
                    *reinterpret_cast<uint2*>(xp) = pack4(dI); *reinterpret_cast<uint2*>(xp + gst) = pack4(dF);
                    *reinterpret_cast<uint2*>(xp + 2 * gst) = pack4(dG); *reinterpret_cast<uint2*>(xp + 3 * gst) = pack4(dO);
                    gi[i] = dI; gf[i] = dF; gg[i] = dG; go[i] = dO;
                    dbacc[c][0].x += dI.x; dbacc[c][0].y += dI.y; dbacc[c][0].z += dI.z; dbacc[c][0].w += dI.w;
                    dbacc[c][1].x += dF.x; dbacc[c][1].y += dF.y; dbacc[c][1].z += dF.z; dbacc[c][1].w += dF.w;
                    dbacc[c][2].x += dG.x; dbacc[c][2].y += dG.y; dbacc[c][2].z += dG.z; dbacc[c][2].w += dG.w;
                    dbacc[c][3].x += dO.x; dbacc[c][3].y += dO.y; dbacc[c][3].z += dO.z; dbacc[c][3].w += dO.w;
                }
                named_bar_sync(1, 128);
                if (te == 0) REC_STAMP(8);
                if (te == 0 && s + 1 < T) red_release_gpu_add(a.ctr + dir * a.nslices + slice, 1u);
                if (te == 0) REC_STAMP(9);
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int b = b0 + rl0 + i;
                    if (b >= a.B) continue;
                    if (!a.dbp) {
                        float* gp = a.gates + ((((long long)b * T + t) * a.ndir + dir) * G4) + u0;
                        *reinterpret_cast<float4*>(gp) = gi[i]; *reinterpret_cast<float4*>(gp + H) = gf[i];
                        *reinterpret_cast<float4*>(gp + 2 * H) = gg[i]; *reinterpret_cast<float4*>(gp + 3 * H) = go[i];
                    }
                    // the (B*T, NG) bf16 copy the dX / dW GEMMs consume
                    __nv_bfloat16* bp = a.dgb + ((long long)b * T + t) * NG + dir * G4 + u0;
                    *reinterpret_cast<uint2*>(bp) = pack4(gi[i]); *reinterpret_cast<uint2*>(bp + H) = pack4(gf[i]);
                    *reinterpret_cast<uint2*>(bp + 2 * H) = pack4(gg[i]); *reinterpret_cast<uint2*>(bp + 3 * H) = pack4(go[i]);
                }
                if (te == 0) REC_STAMP(10);
            }
        }
    }
    if (a.dbp) {
        // bias gradients: the four epilogue warps hold partial sums of the same 32 units over different batch rows; add them up in
        // a fixed order and store this (direction, batch slice)'s row -- the host sums the few slice rows (deterministic)
        __syncthreads();
        float* red = part0;                   // [chain][q][rr][gate][32 units] floats (8 K floats; the partial tiles are idle now)
        if (warp >= 4) {
#pragma unroll
            for (int c = 0; c < MAX_CHAINS; ++c)
#pragma unroll
                for (int gq = 0; gq < 4; ++gq) *reinterpret_cast<float4*>(red + ((((c * 4 + q) * 4 + rr) * 4 + gq) * 32) + 4 * jj) = dbacc[c][gq];
        }
        __syncthreads();
        if (warp == 4) {
            for (int c = 0; c < a.chains; ++c) {
                const int slice = sg + c * a.bsg;
                if (slice >= a.nslices) continue;
#pragma unroll
                for (int gq = 0; gq < 4; ++gq) {
                    float v = 0.f;
#pragma unroll
                    for (int pq = 0; pq < 16; ++pq) v += red[(((c * 16 + pq) * 4 + gq) * 32) + j];      // fixed order over (q, rr)
                    a.dbp[((long long)(dir * a.nslices + slice) * 4 + gq) * H + ub * 128 + kq * 32 + j] = v;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                               // no CTA of the cluster exits while a peer may still read its smem
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// BPTT with the SM-to-SM exchange (default when one chain per CTA): the whole (direction, batch slice) group -- 4 gate CTAs x
// H/128 unit blocks -- is ONE thread-block cluster; the d(pre-activation) slices are pushed into the consumers' operand
// buffers with bulk async DSMEM copies exactly like h_t in lstm_rec_fwd_dsm_kernel (no counter, no release fence, no TMA fetch,
// no global exchange buffer), and the gate-partial reduction uses the same push.  W_hh^T tile resident in tensor memory.
// Cluster rank = ub * 4 + kq.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NTHREADS8, 1) lstm_rec_bwd_dsm_kernel(const RecTcBwdArgs a) {
    constexpr bool WTMEM = true;
    constexpr int NC = 1;              // chains per CTA: this kernel is only launched with one (launch_bwd_dsm); the barrier layout keeps MAX_CHAINS slots
    extern __shared__ uint8_t smem_raw[];
    if (a.start_ctr && threadIdx.x == 0) atomicAdd(a.start_ctr, 1u);   // "this CTA is resident"

    const int H = a.H, T = a.T, KB = H / 64;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int RS = 4 * (H / 128);                                      // CTAs of the group = cluster size
    const uint32_t TILE = (uint32_t)RS * 2048u;                        // 32 rows x H bf16, core-matrix (no-swizzle) layout
    const uint32_t b_sm = base;                                        // [parity][TILE] operand buffers
    const uint32_t xst_sm = b_sm + 2 * TILE;                           // [4 gates][2 KB]: this CTA's dG slices, staged for the push
    uint8_t* xst_ptr = smem_raw + (xst_sm - smem_u32(smem_raw));
    const uint32_t part_off = (xst_sm - smem_u32(smem_raw)) + 4 * 2048;
    // partial-sum exchange inside the 4-CTA cluster, pushed (no cluster barrier, no remote loads): stg[parity][piece q] = this
    // CTA's partial for the 32 units CTA q finalises ([32 rows][32 units] fp32, 4 KB), rcv[parity][src] = what the four CTAs sent me
    float* part0 = reinterpret_cast<float*>(smem_raw + part_off);      // stg: [2][4][32][32]
    float* rcv0 = part0 + 2 * 4 * 1024;                                // rcv: [2][4][32][32]
    const uint32_t part_saddr0 = smem_u32(smem_raw) + part_off;
    const uint32_t rcv_saddr0 = part_saddr0 + 2 * 4 * 4096;
    const uint32_t bar_base = (part_saddr0 + 4 * 4 * 4096 + 15u) & ~15u;
    auto full_bar = [&](int c) { return bar_base + 8u * c; };
    auto tfull_bar = [&](int c) { return bar_base + 8u * (MAX_CHAINS + c); };
    const uint32_t wbar = bar_base + 8u * (2 * MAX_CHAINS);
    const uint32_t tmem_slot = bar_base + 8u * (2 * MAX_CHAINS + 1);
    const int NP = rec_pieces(KB), KBP = KB / NP;                 // the dG tile arrives in pieces (see the forward kernel)
    auto piece_bar = [&](int c, int pc) { return pc == 0 ? full_bar(c) : bar_base + 8u * (2 * MAX_CHAINS + 2 + (pc - 1) * MAX_CHAINS + c); };
    auto rbar = [&](int par) { return bar_base + 8u * (2 * MAX_CHAINS + 2 + 3 * MAX_CHAINS + par); };      // partials received
    auto hbar = [&](int par, int half) { return bar_base + 8u * (2 * MAX_CHAINS + 2 + 3 * MAX_CHAINS + 2 + par * 2 + half); };   // operand halves landed
    const uint32_t tempty = bar_base + 8u * (2 * MAX_CHAINS + 2 + 3 * MAX_CHAINS + 6);                     // accumulators read (epilogue -> MMA)
    const int halfsrc = RS / 2;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int crank = (int)cluster_ctarank();         // == blockIdx.x: the cluster spans the whole group
    const int kq = crank & 3;                         // gate handled by this CTA's reduction slice
    const int ub = crank >> 2;                        // 128-unit block
    const int sg = blockIdx.y, dir = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int F = a.ndir * H, G4 = 4 * H, NG = a.ndir * G4;
    const long long brow = (long long)(T + 2) * F;
    const unsigned group = (unsigned)(4 * (H / 128));  // CTAs sharing one (direction, batch slice)

    if (warp == 0 && lane == 0) {
        for (int c = 0; c < NC; ++c) {
            mbar_init(tfull_bar(c), 1);
            for (int pc = 0; pc < 4; ++pc) mbar_init(piece_bar(c, pc), 1);
        }
        mbar_init(wbar, 1);
        mbar_init(rbar(0), 1); mbar_init(rbar(1), 1);
        for (int i = 0; i < 4; ++i) mbar_init(hbar(i >> 1, i & 1), 1);
        mbar_init(tempty, 8);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // TMEM map: BWD_NACC independent accumulators per chain first, then (WTMEM) the resident W_hh^T tile (H/2 columns)
    const uint32_t tmem_cols = WTMEM ? 512u : 256u;
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    const uint32_t tmem_w = tmem_base + MAX_CHAINS * BWD_NACC * NB_SLICE;
    if (WTMEM) {
        if (warp >= 4 && warp < 8) {
            // A operand in tensor memory: TMEM lane = unit row of the block, column pair = two consecutive k (gate rows of gate kq)
            const int qq = warp & 3;
            const uint32_t* wrow = reinterpret_cast<const uint32_t*>(a.w_t + ((long long)(dir * H + ub * 128 + qq * 32 + lane)) * G4 + kq * H);
            for (int cb = 0; cb < H / 64; ++cb) {
                uint32_t v[32];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const uint4 t4 = *reinterpret_cast<const uint4*>(wrow + cb * 32 + i * 4);
                    v[i * 4 + 0] = t4.x; v[i * 4 + 1] = t4.y; v[i * 4 + 2] = t4.z; v[i * 4 + 3] = t4.w;
                }
                tmem_st32(tmem_w + ((uint32_t)(qq * 32) << 16) + cb * 32, v);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }

    cluster_sync_all();              // every CTA of the cluster has initialised its barriers before a peer pushes partials at it

    // per-thread state of the epilogue role (declared for all so the step loop below is shared by every warp)
    // Epilogue ownership: thread (warp q, lane) finalises FOUR consecutive units u0..u0+3 (u0 = block base + 4*(lane & 7)) for TWO
    // batch rows (q*8 + 2*(lane >> 3) + {0, 1}), so that every global access of the pointwise backward is a 128-bit one: a quarter
    // of the memory instructions of a one-unit-per-lane mapping (their issue rate, not the exchange, had become the critical path).
    // EIGHT epilogue warps (round 2; 4 units x ONE batch row per thread): warp e = warp - 4 reads TMEM lane quarter q = e & 3, batch columns
    // 16 (e >> 2) .. + 16 of the accumulators, and finalises batch rows 4e .. 4e + 3 of the slice
    const int q = warp & 3, j = lane;
    const int e8 = warp >= 4 ? warp - 4 : 0, hf8 = e8 >> 2;
    const int te = (warp - 4) * 32 + lane;
    const int jj = lane & 7, rr = lane >> 3;
    const int u0 = ub * 128 + kq * 32 + 4 * jj;        // first of the 4 units this thread finalises
    const int rl0 = e8 * 4 + rr;                       // its batch row inside the 32-row slice
    constexpr int NR = 1;                              // rows per thread
    float4 dcst[NC][NR];
    float4 mkr[NC][NR];                        // locked-dropout mask: constant over time
    float4 cnext[NC][NR];                      // c_{t-1} loaded at this step = c_t of the next step (time runs backwards)
    float4 dbacc[NC][4];                       // bias-gradient partial sums (gate x 4 units) over this thread's rows, all steps
    int lenr[NC][NR];
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int c = 0; c < NC; ++c) {
#pragma unroll
        for (int gq = 0; gq < 4; ++gq) dbacc[c][gq] = z4;
#pragma unroll
        for (int i = 0; i < NR; ++i) {
            dcst[c][i] = z4; cnext[c][i] = z4;
            const int b = (sg + c * a.bsg) * NB_SLICE + rl0 + i;
            lenr[c][i] = (warp >= 4 && c < a.chains && sg + c * a.bsg < a.nslices && b < a.B) ? a.lens[b] : 0;
            mkr[c][i] = (a.mask && lenr[c][i] > 0) ? *reinterpret_cast<const float4*>(a.mask + (long long)b * F + dir * H + u0)
                                                   : make_float4(1.f, 1.f, 1.f, 1.f);
        }
    }

    // Every warp walks the same (step, chain) sequence: the cluster barrier of each iteration needs all threads of all 4 CTAs.
    int iter = 0, riter = 0;
    for (int s = 0; s < T; ++s) {
        const int t = (dir == 0) ? (T - 1 - s) : s;
        const int t_prev = (dir == 0) ? (T - s) : (s - 1);
        const int fprev = (dir == 0) ? t : t + 2, fcur = t + 1;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const int slice = sg + c * a.bsg;
            if (c >= a.chains || slice >= a.nslices) continue;          // uniform across the cluster (same sg, chains)
            const int b0 = slice * NB_SLICE;
            const int rpar = riter & 1;                                   // parity of this reduce-iteration (s > 0 only)
            const uint32_t rphase = (uint32_t)((riter >> 1) & 1);
            if (s > 0) ++riter;
            ++iter;
            float4 gi[NR], gf[NR], gg[NR], go[NR], ct[NR], cp[NR], dh[NR], rec[NR];
            bool valid[NR];
            if (warp == 0) {
                // (no producer: the operand is pushed into this CTA's buffers by its peers)
            } else if (warp == 1) {
                if (s > 0) {
                    // whole warp waits, one elected lane issues; BWD_NACC independent accumulators (one per k sub-step)
                    const int par = (s - 1) & 1;                               // buffer holding dG of step s-1
                    const uint32_t phase = (uint32_t)(((s - 1) >> 1) & 1);
                    const uint32_t buf = b_sm + (uint32_t)par * TILE;
                    if (s > 1) mbar_wait(tempty, (uint32_t)(s & 1));           // the epilogue has read step s-1's accumulators
                    for (int hf = 0; hf < 2; ++hf) {
                        if (lane == 0) mbar_arrive_expect_tx(hbar(par, hf), (uint32_t)halfsrc * 2048u);
                        __syncwarp();
                        mbar_wait(hbar(par, hf), phase);
                        if (lane == 0 && hf == 0) REC_STAMP(2);
                        tc_fence_after();
                        if (elect_one()) {
                            const int ks_lo = hf * halfsrc * 2, ks_hi = ks_lo + halfsrc * 2;      // 16-unit k-steps
                            for (int ks = ks_lo; ks < ks_hi; ++ks) {
                                const uint32_t d_tmem = tmem_base + (uint32_t)((c * BWD_NACC + (ks & 3)) * NB_SLICE);
                                umma_bf16_ts(d_tmem, tmem_w + ks * 8, make_desc_k_noswz(buf + (uint32_t)ks * 1024u), IDESC, ks >= 4 ? 1u : 0u);
                            }
                            if (hf == 1) umma_commit(tfull_bar(c));
                        }
                        __syncwarp();
                    }
                    if (lane == 0) REC_STAMP(3);
                }
                __syncwarp();
            } else if (warp >= 4) {
                // operands of the pointwise backward for this thread's unit, issued before waiting on the tensor pipe
#pragma unroll
                for (int i = 0; i < NR; ++i) {
                    const int b = b0 + rl0 + i;
                    valid[i] = t < lenr[c][i];
                    // pure 128-bit loads, no branches: rows past B are clamped to a valid address; invalid rows are zeroed in the
                    // pointwise step
                    const int bc = b < a.B ? b : a.B - 1;
                    const float* gp = a.gates + (((long long)bc * T + t) * a.ndir + dir) * G4 + u0;
                    gi[i] = *reinterpret_cast<const float4*>(gp); gf[i] = *reinterpret_cast<const float4*>(gp + H);
                    gg[i] = *reinterpret_cast<const float4*>(gp + 2 * H); go[i] = *reinterpret_cast<const float4*>(gp + 3 * H);
                    // c_t: carried over from the previous step's c_{t-1} load (first step: loaded)
                    if (s == 0) ct[i] = *reinterpret_cast<const float4*>(a.cs_pad + (long long)bc * brow + (long long)fcur * F + dir * H + u0);
                    else ct[i] = cnext[c][i];
                    cp[i] = *reinterpret_cast<const float4*>(a.cs_pad + (long long)bc * brow + (long long)fprev * F + dir * H + u0);
                    dh[i] = *reinterpret_cast<const float4*>(a.dout + ((long long)bc * T + t) * F + dir * H + u0);
                    rec[i] = z4;
                }
                if (te == 0) REC_STAMP(4);
                if (s > 0) {
                    mbar_wait(tfull_bar(c), (uint32_t)((s - 1) & 1));
                    if (te == 0) REC_STAMP(5);
                    tc_fence_after();
                    float accv[16];
#pragma unroll
                    for (int acc = 0; acc < BWD_NACC; ++acc) {
                        uint32_t v[16];
                        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((c * BWD_NACC + acc) * NB_SLICE + 16 * hf8), v);
#pragma unroll
                        for (int n = 0; n < 16; ++n) accv[n] = acc ? accv[n] + __uint_as_float(v[n]) : __uint_as_float(v[n]);
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tempty);
                    // TMEM lane = unit (32q + lane) of the block, column = batch row.  Warps q and 4+q hold the 32 units CTA q of the
                    // cluster finalises, 16 batch rows each: stage them as [row][unit] and push the 2 KB half piece into CTA q's receive buffer
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the piece staged two iterations ago has been read
                    __syncwarp();
                    float* stg = part0 + (rpar * 4 + q) * 1024 + hf8 * 512;
#pragma unroll
                    for (int n = 0; n < 16; ++n) stg[n * 32 + lane] = accv[n];
                    tc_fence_before();
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) {
                        const uint32_t peer = (uint32_t)(ub * 4 + q);
                        const uint32_t dst = mapa_u32(rcv_saddr0 + (uint32_t)((rpar * 4 + kq) * 4096 + hf8 * 2048), peer);
                        bulk_copy_to_peer(dst, part_saddr0 + (uint32_t)((rpar * 4 + q) * 4096 + hf8 * 2048), 2048u, mapa_u32(rbar(rpar), peer));
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                    if (te == 0) {
                        mbar_arrive_expect_tx(rbar(rpar), 4u * 4096u);         // the four pieces for my 32 units
                        REC_STAMP(6);
                    }
                }
            }
            if (s > 0) {
                if (warp >= 4) {
                    mbar_wait(rbar(rpar), rphase);
                    if (te == 0) REC_STAMP(7);
                    const float* rcv = rcv0 + rpar * 4 * 1024;
#pragma unroll
                    for (int i = 0; i < NR; ++i) {
                        const int o = (rl0 + i) * 32 + 4 * jj;
                        const float4 p0 = *reinterpret_cast<const float4*>(rcv + o), p1 = *reinterpret_cast<const float4*>(rcv + 1024 + o);
                        const float4 p2 = *reinterpret_cast<const float4*>(rcv + 2048 + o), p3 = *reinterpret_cast<const float4*>(rcv + 3072 + o);
                        rec[i] = make_float4((p0.x + p1.x) + (p2.x + p3.x), (p0.y + p1.y) + (p2.y + p3.y), (p0.z + p1.z) + (p2.z + p3.z),
                                             (p0.w + p1.w) + (p2.w + p3.w));
                    }
                }
                if (warp == 4 && lane == 0) REC_STAMP(11);
            }
            if (warp >= 4) {
                // the dG slices staged at the previous step must have been read by their bulk copies before they are overwritten.
                // (Thread te < RS also issued a partial-piece copy this iteration when lane == 0: wait for everything but that one.)
                if (te < RS && s > 0) {
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                }
                named_bar_sync(1, 256);
                auto pw = [&](float gi_, float gf_, float gg_, float go_, float ct_, float cp_, float dh_, float mk_, float rec_, float dc_,
                              bool ok, float& dai, float& daf, float& dag, float& dao, float& dcn) {
                    dai = daf = dag = dao = dcn = 0.f;
                    if (ok) {
                        const float tcv = tanh_fast(ct_);
                        const float dhv = fmaf(dh_, mk_, rec_);
                        const float dct = fmaf(dhv * go_, 1.f - tcv * tcv, dc_);
                        dai = dct * gg_ * gi_ * (1.f - gi_);
                        daf = dct * cp_ * gf_ * (1.f - gf_);
                        dag = dct * gi_ * (1.f - gg_ * gg_);
                        dao = dhv * tcv * go_ * (1.f - go_);
                        dcn = dct * gf_;
                    }
                };
                auto pack4 = [](const float4& v) {
                    const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
                    uint2 r;
                    r.x = *reinterpret_cast<const uint32_t*>(&lo); r.y = *reinterpret_cast<const uint32_t*>(&hi);
                    return r;
                };
#pragma unroll
                for (int i = 0; i < NR; ++i) {
                    const int b = b0 + rl0 + i;
                    float4 dI, dF, dG, dO, dC;
                    pw(gi[i].x, gf[i].x, gg[i].x, go[i].x, ct[i].x, cp[i].x, dh[i].x, mkr[c][i].x, rec[i].x, dcst[c][i].x, valid[i], dI.x, dF.x, dG.x, dO.x, dC.x);
                    pw(gi[i].y, gf[i].y, gg[i].y, go[i].y, ct[i].y, cp[i].y, dh[i].y, mkr[c][i].y, rec[i].y, dcst[c][i].y, valid[i], dI.y, dF.y, dG.y, dO.y, dC.y);
                    pw(gi[i].z, gf[i].z, gg[i].z, go[i].z, ct[i].z, cp[i].z, dh[i].z, mkr[c][i].z, rec[i].z, dcst[c][i].z, valid[i], dI.z, dF.z, dG.z, dO.z, dC.z);
                    pw(gi[i].w, gf[i].w, gg[i].w, go[i].w, ct[i].w, cp[i].w, dh[i].w, mkr[c][i].w, rec[i].w, dcst[c][i].w, valid[i], dI.w, dF.w, dG.w, dO.w, dC.w);
                    dcst[c][i] = dC;
                    cnext[c][i] = cp[i];
                    // what the peers' next step reads: bf16 d(pre-activation), staged per gate in the UMMA no-swizzle core-matrix layout
                    // (core (kc = unit/8, ng = row/8) at kc*512 + ng*128; row-in-core 16 B apart): 4 units = one 64-bit store
                    if (s + 1 < T) {
                        uint8_t* xp = xst_ptr + (jj >> 1) * 512 + (e8 >> 1) * 128 + ((e8 & 1) * 4 + rr + i) * 16 + (jj & 1) * 8;
                        *reinterpret_cast<uint2*>(xp) = pack4(dI); *reinterpret_cast<uint2*>(xp + 2048) = pack4(dF);
                        *reinterpret_cast<uint2*>(xp + 4096) = pack4(dG); *reinterpret_cast<uint2*>(xp + 6144) = pack4(dO);
                    }
                    gi[i] = dI; gf[i] = dF; gg[i] = dG; go[i] = dO;
                    dbacc[c][0].x += dI.x; dbacc[c][0].y += dI.y; dbacc[c][0].z += dI.z; dbacc[c][0].w += dI.w;
                    dbacc[c][1].x += dF.x; dbacc[c][1].y += dF.y; dbacc[c][1].z += dF.z; dbacc[c][1].w += dF.w;
                    dbacc[c][2].x += dG.x; dbacc[c][2].y += dG.y; dbacc[c][2].z += dG.z; dbacc[c][2].w += dG.w;
                    dbacc[c][3].x += dO.x; dbacc[c][3].y += dO.y; dbacc[c][3].z += dO.z; dbacc[c][3].w += dO.w;
                }
                if (s + 1 < T) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                named_bar_sync(1, 256);
                if (te == 0) REC_STAMP(8);
                if (s + 1 < T && te < RS) {
                    // thread `te` pushes the slice of gate g = te & 3 into the operand buffer of the consumer of that gate in unit block
                    // te >> 2 (cluster rank (te >> 2) * 4 + g), at this CTA's position, and signals the half its rank belongs to
                    const int g = te & 3, par = s & 1;
                    const uint32_t peer = (uint32_t)((te >> 2) * 4 + g);
                    const uint32_t dst = mapa_u32(b_sm + (uint32_t)par * TILE + (uint32_t)crank * 2048u, peer);
                    bulk_copy_to_peer(dst, xst_sm + (uint32_t)g * 2048u, 2048u, mapa_u32(hbar(par, crank >= halfsrc ? 1 : 0), peer));
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                if (te == 0) REC_STAMP(9);
#pragma unroll
                for (int i = 0; i < NR; ++i) {
                    const int b = b0 + rl0 + i;
                    if (b >= a.B) continue;
                    if (!a.dbp) {
                        float* gp = a.gates + ((((long long)b * T + t) * a.ndir + dir) * G4) + u0;
                        *reinterpret_cast<float4*>(gp) = gi[i]; *reinterpret_cast<float4*>(gp + H) = gf[i];
                        *reinterpret_cast<float4*>(gp + 2 * H) = gg[i]; *reinterpret_cast<float4*>(gp + 3 * H) = go[i];
                    }
                    // the (B*T, NG) bf16 copy the dX / dW GEMMs consume
                    __nv_bfloat16* bp = a.dgb + ((long long)b * T + t) * NG + dir * G4 + u0;
                    *reinterpret_cast<uint2*>(bp) = pack4(gi[i]); *reinterpret_cast<uint2*>(bp + H) = pack4(gf[i]);
                    *reinterpret_cast<uint2*>(bp + 2 * H) = pack4(gg[i]); *reinterpret_cast<uint2*>(bp + 3 * H) = pack4(go[i]);
                }
                if (te == 0) REC_STAMP(10);
            }
        }
        // progress for a consumer on another stream (the dX GEMM of this layer, tile by tile): the epilogue warps have stored the bf16
        // gate gradients of step s; warp 0 (idle in this kernel) makes them visible device-wide and counts the CTA in
        if (a.progress && (s + 1) % a.progress_every == 0) {
            if (warp >= 4) {
                asm volatile("bar.arrive 5, 288;" ::: "memory");
            } else if (warp == 0) {
                asm volatile("bar.sync 5, 288;" ::: "memory");
                if (lane == 0) {
                    __threadfence();
                    atomicAdd(a.progress + dir * gridDim.y + sg, 1u);
                }
            }
        }
    }
    if (a.dbp) {
        // bias gradients: the four epilogue warps hold partial sums of the same 32 units over different batch rows; add them up in
        // a fixed order and store this (direction, batch slice)'s row -- the host sums the few slice rows (deterministic)
        __syncthreads();
        float* red = part0;                   // [chain][warp e][rr][gate][32 units] floats (8 K floats; the partial tiles are idle now)
        if (warp >= 4) {
#pragma unroll
            for (int c = 0; c < NC; ++c)
#pragma unroll
                for (int gq = 0; gq < 4; ++gq) *reinterpret_cast<float4*>(red + ((((c * 8 + e8) * 4 + rr) * 4 + gq) * 32) + 4 * jj) = dbacc[c][gq];
        }
        __syncthreads();
        if (warp == 4) {
            for (int c = 0; c < a.chains; ++c) {
                const int slice = sg + c * a.bsg;
                if (slice >= a.nslices) continue;
#pragma unroll
                for (int gq = 0; gq < 4; ++gq) {
                    float v = 0.f;
#pragma unroll
                    for (int pq = 0; pq < 32; ++pq) v += red[(((c * 32 + pq) * 4 + gq) * 32) + j];      // fixed order over (warp, rr)
                    a.dbp[((long long)(dir * a.nslices + slice) * 4 + gq) * H + ub * 128 + kq * 32 + j] = v;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                               // no CTA of the cluster exits while a peer may still read its smem
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}

}  // namespace

// BPTT with the SM-to-SM exchange; returns LAS_ERR_UNSUPPORTED when the shape / device does not allow it (caller falls back)
static int launch_bwd_dsm(const float* dout, float* gates, void* dgates_bf16, const float* cs_pad, const void* w_hh_t_bf16, const int* lens,
                          const float* drop_mask, int B, int T, int H, int ndir, cudaStream_t st, void* stream, float* dbp) {
    const char* de = getenv("LAS_REC_DSMEM");
    if (de && atoi(de) == 0) return LAS_ERR_UNSUPPORTED;
    if (H % 128 != 0 || H > 512) return LAS_ERR_UNSUPPORTED;
    const LasDeviceInfo* di = las_device_info();
    const int rs = 4 * (H / 128);
    const int nslices = ceil_div(B, NB_SLICE);
    const int max_bsg = di->num_sms / (rs * ndir);
    if (max_bsg < 1 || nslices > max_bsg) return LAS_ERR_UNSUPPORTED;          // one chain per CTA only
    if (nslices * ndir > 6 && !(de && atoi(de) == 2)) return LAS_ERR_UNSUPPORTED;  // 8 clusters at once exchange at half the speed (see the forward launcher)
    const size_t smem = 1024 + 2 * (size_t)rs * 2048 + 4 * 2048 + 4 * 4 * 4096 + 16 + 8 * (2 * MAX_CHAINS + 2 + 3 * MAX_CHAINS + 7) + 64;
    if (smem > (size_t)di->max_smem_optin) return LAS_ERR_UNSUPPORTED;
    RecTcBwdArgs a{};
    a.gates = gates; a.dgb = (__nv_bfloat16*)dgates_bf16; a.dout = dout; a.cs_pad = cs_pad; a.lens = lens; a.mask = drop_mask;
    a.B = B; a.T = T; a.H = H; a.ndir = ndir; a.nslices = nslices; a.chains = 1; a.bsg = nslices; a.KBr = 4 * H / 64; a.CH = H / 64;
    a.dbg = g_rec_dbg; a.Bpad = nslices * NB_SLICE;
    a.w_t = (const __nv_bfloat16*)w_hh_t_bf16;
    a.dbp = dbp;
    if (t_bprog_ctr && t_bprog_every > 0 && nslices * ndir <= 64) { a.progress = t_bprog_ctr; a.progress_every = t_bprog_every; }
    auto kd = lstm_rec_bwd_dsm_kernel;
    if (cudaFuncSetAttribute(kd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return LAS_ERR_UNSUPPORTED; }
    if (rs > 8 && cudaFuncSetAttribute(kd, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) { cudaGetLastError(); return LAS_ERR_UNSUPPORTED; }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(rs, nslices, ndir); cfg.blockDim = dim3(NTHREADS8); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = rs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int nclusters = 0;
    if (cudaOccupancyMaxActiveClusters(&nclusters, kd, &cfg) != cudaSuccess || nclusters < 1) { cudaGetLastError(); return LAS_ERR_UNSUPPORTED; }
    LasProfScope prof(LAS_PROF_REC_BWD, stream, (double)T);
    int dev = -1;
    unsigned target = 0;
    if (t_start_armed && nclusters >= nslices * ndir && cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64) {
        std::lock_guard<std::mutex> lk(g_start_mu);
        if (!g_start_ctr[dev]) {
            if (cudaMalloc(&g_start_ctr[dev], 256) != cudaSuccess || cudaMemset(g_start_ctr[dev], 0, 256) != cudaSuccess) {
                cudaGetLastError();
                g_start_ctr[dev] = nullptr;
            }
            g_start_target[dev] = 0;
        }
        if (g_start_ctr[dev]) {
            a.start_ctr = g_start_ctr[dev];
            g_start_target[dev] += (unsigned)(rs * nslices * ndir);
            target = g_start_target[dev];
        }
    }
    if (cudaLaunchKernelEx(&cfg, kd, a) != cudaSuccess) {
        cudaGetLastError();
        if (a.start_ctr) {          // no CTA ran: take the target back so that counter and target stay in step
            std::lock_guard<std::mutex> lk(g_start_mu);
            g_start_target[dev] -= (unsigned)(rs * nslices * ndir);
        }
        return LAS_ERR_UNSUPPORTED;
    }
    if (a.start_ctr) {
        WaitValue32Fn wait_value32 = get_wait_value32();
        if (wait_value32 && wait_value32((CUstream)t_start_stream, (CUdeviceptr)a.start_ctr, target, CU_STREAM_WAIT_VALUE_GEQ) == CUDA_SUCCESS) {
            t_start_armed = false;
            t_start_mode = 1;
        }                           // else: still armed -> the caller's scope makes the side stream wait for completion
    }
    if (a.progress) { t_prog_clusters = nslices * ndir; t_prog_rs = rs; }
    las_count_launch(1);
    return LAS_OK;
}

// returns LAS_OK, or a negative code when this variant cannot run (caller falls back to las_lstm_rec_bwd_tc's streaming variant)
static int launch_bwd_tc2(const float* dout, float* gates, void* dgates_bf16, const float* cs_pad, const void* w_hh_t_bf16, const int* lens,
                          const float* drop_mask, int B, int T, int H, int ndir, void* ws, cudaStream_t st, void* stream, float* dbp) {
    if (H % 128 != 0) return LAS_ERR_UNSUPPORTED;
    const LasDeviceInfo* di = las_device_info();
    const int rs = 4 * (H / 128);
    const int nslices = ceil_div(B, NB_SLICE);
    int max_bsg = di->num_sms / (rs * ndir);
    if (max_bsg < 1) return LAS_ERR_UNSUPPORTED;
    const int bsg = nslices < max_bsg ? nslices : max_bsg;
    const int chains = ceil_div(nslices, bsg);
    if (chains > MAX_CHAINS) return LAS_ERR_UNSUPPORTED;
    const int KB = H / 64;
    const char* wt_env = getenv("LAS_REC_WTMEM");
    const bool wtmem = (wt_env ? atoi(wt_env) != 0 : true) && H <= 512;
    const size_t smem = 1024 + (wtmem ? 0 : (size_t)KB * 16384) + (size_t)chains * KB * 4096 + 4 * 4 * 4096 + 16 + 8 * (2 * MAX_CHAINS + 2 + 3 * MAX_CHAINS + 2) + 64;
    if (smem > (size_t)di->max_smem_optin) return LAS_ERR_UNSUPPORTED;
    RecTcBwdArgs a{};
    a.w_t = (const __nv_bfloat16*)w_hh_t_bf16;
    a.dbp = dbp;
    a.gates = gates; a.dgb = (__nv_bfloat16*)dgates_bf16; a.dout = dout; a.cs_pad = cs_pad; a.lens = lens; a.mask = drop_mask;
    a.ctr = (unsigned*)ws;
    a.B = B; a.T = T; a.H = H; a.ndir = ndir; a.nslices = nslices; a.chains = chains; a.bsg = bsg; a.KBr = 4 * H / 64; a.CH = KB; a.dbg = g_rec_dbg;
    const long long NG = (long long)ndir * 4 * H;
    CUtensorMap tmWt, tmG;
    int rc = make_map_2d(&tmWt, w_hh_t_bf16, 4LL * H, (long long)ndir * H, 64, 128);
    if (rc) return rc;
    a.Bpad = nslices * NB_SLICE;
    a.dgx = (__nv_bfloat16*)((char*)ws + 1024);
    {   // exchange buffer (ndir*2*4*Bpad rows, H) viewed as (64, rows, H/64): one box = a 32 x H tile in k-block-major smem order
        const int KBPt = KB / rec_pieces(KB);
        const long long dims[3] = {64, (long long)ndir * 2 * 4 * a.Bpad, KB};
        const long long strides[2] = {H, 64};
        const int box[3] = {64, NB_SLICE, KBPt};      // one piece of the tile per TMA issue
        rc = make_map_nd(&tmG, a.dgx, 3, dims, strides, box);
        if (rc) return rc;
    }
    (void)NG;
    void* kern = wtmem ? (void*)lstm_rec_bwd_tc2_kernel<true> : (void*)lstm_rec_bwd_tc2_kernel<false>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        cudaGetLastError();
        return LAS_ERR_UNSUPPORTED;
    }
    LAS_CUDA(cudaMemsetAsync(ws, 0, 1024, st));
    LasProfScope prof(LAS_PROF_REC_BWD, stream, (double)T);
    void* args[] = {(void*)&tmWt, (void*)&tmG, (void*)&a};
    cudaError_t e = cudaLaunchCooperativeKernel(kern, dim3(rs, bsg, ndir), dim3(NTHREADS), args, smem, st);
    if (e != cudaSuccess) {
        las_set_error("lstm_rec_bwd_tc2 launch failed: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return LAS_ERR_UNSUPPORTED;
    }
    las_count_launch(1);
    return LAS_OK;
}
