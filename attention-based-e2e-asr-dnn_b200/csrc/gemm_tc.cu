// bf16 tensor-core GEMM for the all-timestep LSTM gate projections: tcgen05.mma + TMEM accumulators, TMA-fed,
// warp-specialised and persistent (one CTA per SM).  sm_100a only.
//
// Replaces the input half of nn.LSTM (X . W_ih^T for every timestep at once, reference src/modules.py:80,189) and its
// two autograd GEMMs (dX = dG . W_ih, dW = dG^T . X).  The pyramidal frame-pair concat / odd-frame drop
// (src/modules.py:171-185), the "T = max(lx)" truncation and the batch padding are folded into the TMA tensor maps:
// every operand is described as a 3-D tensor (contiguous dim, row dim, batch dim) with arbitrary row / batch strides,
// and out-of-range rows are zero-filled by the TMA unit, so no reshaped / packed copy of the activations ever exists.
//
// Tile: 128 (M) x 256 (N) x 64 (K) per stage, 4 stages of 48 KB, UMMA 128x256x16 (cta_group::1), fp32 accumulators
// double-buffered in TMEM (2 x 256 columns) so the epilogue of tile i overlaps the MMAs of tile i+1.
//   warp 0 : TMA producer (one elected lane)        warp 1 : MMA issuer (one elected lane)
//   warp 2 : TMEM allocator                          warps 4-7 : epilogue (tcgen05.ld -> +bias -> global)
// Operand layouts (SWIZZLE_128B everywhere):
//   K-major  operand tile rows x 64k : one TMA box {64, rows, 1}; UMMA desc SBO = 1024 B, K-advance = 32 B
//   MN-major operand tile 64k x mn   : mn/64 TMA boxes {64, 64, 1} of 8 KB; UMMA desc LBO = 8 KB, SBO = 1024 B,
//                                      K-advance = 2048 B
#include "las_common.cuh"
#include <cuda_fp16.h>
#include <stdlib.h>
#include "las_b200.h"
#include <cuda.h>
#include <new>

namespace {

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    return pred != 0;
}

// BN is a template parameter: 256 for the big gate GEMMs (one tile per SM per wave), 64 for the small decoder GEMMs
// (M = batch <= 128): four times as many CTAs, a quarter of the MMA + epilogue time per CTA.
constexpr int BM = 128, BK = 64;
constexpr int STAGES = 4;
constexpr int A_STAGE_BYTES = BM * BK * 2;       // 16 KB
constexpr int EPI_LD = 36;                          // floats per staged row (144 B: 16-byte aligned, conflict-free STS.128/LDS.128)
constexpr int EPI_BIAS = 256;                        // per epilogue warp: bias1 + bias2 of the tile's (at most 256) columns
constexpr int EPI_BYTES = 4 * 32 * EPI_LD * 4 + 4 * EPI_BIAS * 4;      // one 32x32 fp32 transpose tile per epilogue warp + its bias sums
constexpr int NTHREADS = 256;
template <int BN> struct Cfg {
    static constexpr int B_STAGE_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + EPI_BYTES;
    static constexpr int TMEM_COLS = 2 * BN;       // 512 or 128: a power of two >= 32
};

struct TcArgs {
    float* C;
    const float* bias1;
    const float* bias2;
    // output row mapping: tile row r of batch b -> C + b*c_bs + r*ldc ; rows >= R are skipped
    long long c_bs, ldc;
    int R;            // rows per batch (M direction) for K-major A; total M for MN-major A
    int NB;           // batches in the M direction (1 for MN-major A)
    int N;            // output columns
    int mt_per_b;     // M tiles per batch
    int nt;           // N tiles
    int kt_per_b;     // K iterations per K-batch
    int KB;           // batches in the K direction (wgrad: B of (b,t) rows; else 1)
    int accumulate;   // C += result
    const int* lens;  // optional (NB): skip M tiles whose first row >= lens[b] (rows past a sequence's length)
    int b_first;      // first M-batch index (prepared plans address one batch of a multi-batch tensor map per launch)
    LasLstmEpi le;    // fused LSTM-cell epilogue (EPI == 1 instantiation only)
    int splitk;       // >1: the K range of every output tile is split over `splitk` CTAs writing partials to Cpart
    float* Cpart;     // (splitk, R, ldp) partial sums
    long long ldp;
    uint32_t fmt_clear;   // instruction-descriptor format bits to clear: bit 7 -> A is fp16, bit 10 -> B is fp16 (set = bf16)
    // K-major A and B only: the reduction runs over chunks of `kc_iters` K iterations whose starts lie `kc_stride` elements apart in BOTH
    // operands (0: one contiguous K range).  This is how one direction's half of a BiLSTM output feeds the next layer's gate GEMM.
    int kc_iters, kc_stride;
};

// ---- PTX wrappers --------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }

// the load without its wait: the caller overlaps it with other work and calls tmem_ld_wait() before it touches r[]
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

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

// UMMA shared-memory descriptor (cute::UMMA::SmemDescriptor bit layout): start addr [0,14) >>4, LBO [16,30) >>4,
// SBO [32,46) >>4, version [46,48) = 1, layout type [61,64) = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 [4,6)=1, a/b format BF16 [7,10)/[10,13)=1,
// a_major bit 15, b_major bit 16, n_dim [17,23) = N>>3, m_dim [24,29) = M>>4
__host__ __device__ constexpr uint32_t make_idesc(bool a_mn, bool b_mn, int bn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(bn >> 3) << 17) |
           ((uint32_t)(BM >> 4) << 24);
}

__device__ __forceinline__ float tc_tanh_fast(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float tc_sigmoid_fast(float x) { return fmaf(0.5f, tc_tanh_fast(0.5f * x), 0.5f); }

template <bool A_MN, bool B_MN, int BN, int EPI>
__global__ void __launch_bounds__(NTHREADS, 1) gemm_bf16_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                   const __grid_constant__ CUtensorMap tmB, const TcArgs g) {
    constexpr int STAGE_BYTES = Cfg<BN>::STAGE_BYTES;
    constexpr int TMEM_COLS = Cfg<BN>::TMEM_COLS;
    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B operands need 1024-byte aligned stage bases
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = base + STAGES * STAGE_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    // PDL: everything above (barriers, TMEM, descriptor prefetch) may overlap the previous kernel; nothing below may
    pdl_wait();
    pdl_trigger();

    const int splitk = g.splitk > 1 ? g.splitk : 1;
    const int total_tiles = g.NB * g.mt_per_b * g.nt * splitk;
    const int kiters_all = g.KB * g.kt_per_b;
    float* epi_sm = reinterpret_cast<float*>(smem_raw + (bar_base + 256 - smem_u32(smem_raw)));

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer =====
            int stage = 0; uint32_t phase = 0;
            for (int tile0 = blockIdx.x; tile0 < total_tiles; tile0 += gridDim.x) {
                const int split = tile0 % splitk, tile = tile0 / splitk;
                const int ntile = tile % g.nt, mrem = tile / g.nt;
                const int mtile = mrem % g.mt_per_b, b = mrem / g.mt_per_b + g.b_first;
                if (g.lens && mtile * BM >= g.lens[b]) continue;
                const int k_lo = (int)((long long)kiters_all * split / splitk), k_hi = (int)((long long)kiters_all * (split + 1) / splitk);
                for (int kit = k_lo; kit < k_hi; ++kit) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    const uint32_t sa = base + stage * STAGE_BYTES, sb = sa + A_STAGE_BYTES;
                    mbar_arrive_expect_tx(full_bar(stage), STAGE_BYTES);
                    const int kb = kit / g.kt_per_b, kk = kit - kb * g.kt_per_b;
                    const int kcoord = g.kc_iters > 0 ? (kk / g.kc_iters) * g.kc_stride + (kk % g.kc_iters) * BK : kk * BK;
                    if (!A_MN) {
                        tma_load_3d(sa, &tmA, full_bar(stage), kcoord, mtile * BM, b);
                    } else {
#pragma unroll
                        for (int j = 0; j < BM / 64; ++j)
                            tma_load_3d(sa + j * 8192, &tmA, full_bar(stage), mtile * BM + j * 64, kk * BK, kb);
                    }
                    if (!B_MN) {
                        tma_load_3d(sb, &tmB, full_bar(stage), kcoord, ntile * BN, 0);
                    } else {
#pragma unroll
                        for (int j = 0; j < BN / 64; ++j)
                            tma_load_3d(sb + j * 8192, &tmB, full_bar(stage), ntile * BN + j * 64, kk * BK, kb);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: the whole warp walks the loop (waits included) and ONE elected lane issues -- with `if (lane == 0)`
        // the branch is divergent, every tcgen05.mma gets wrapped in a divergence loop with R2UR moves (~65 cycles per issue,
        // more than a 128x64x16 UMMA takes); warp-uniform control flow keeps the descriptors in uniform registers =====
        const uint32_t idesc = make_idesc(A_MN, B_MN, BN) & ~g.fmt_clear;     // kind::f16 takes fp16 / bf16 per operand
        int stage = 0; uint32_t phase = 0;
        int acc = 0; uint32_t accphase = 0;
        for (int tile0 = blockIdx.x; tile0 < total_tiles; tile0 += gridDim.x) {
            const int split = tile0 % splitk, tile = tile0 / splitk;
            const int mrem = tile / g.nt;
            const int mtile = mrem % g.mt_per_b, b = mrem / g.mt_per_b + g.b_first;
            if (g.lens && mtile * BM >= g.lens[b]) continue;
            const int k_lo = (int)((long long)kiters_all * split / splitk), k_hi = (int)((long long)kiters_all * (split + 1) / splitk);
            mbar_wait(tempty_bar(acc), accphase ^ 1u);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * BN;
            for (int kit = k_lo; kit < k_hi; ++kit) {
                mbar_wait(full_bar(stage), phase);
                tc_fence_after();
                const uint32_t sa = base + stage * STAGE_BYTES, sb = sa + A_STAGE_BYTES;
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        const uint64_t ad = A_MN ? make_desc(sa + k * 2048, 8192, 1024) : make_desc(sa + k * 32, 16, 1024);
                        const uint64_t bd = B_MN ? make_desc(sb + k * 2048, 8192, 1024) : make_desc(sb + k * 32, 16, 1024);
                        umma_bf16(d_tmem, ad, bd, idesc, ((kit - k_lo) | k) ? 1u : 0u);
                    }
                    umma_commit(empty_bar(stage));          // frees the smem slot when these MMAs retire
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
            if (elect_one()) umma_commit(tfull_bar(acc));   // accumulator ready for the epilogue
            __syncwarp();
            if (++acc == 2) { acc = 0; accphase ^= 1u; }
        }
    } else if (warp >= 4) {
        // ===== epilogue: TMEM -> registers -> smem transpose -> (+bias, +C) -> 128-byte coalesced global stores =====
        const int q = warp & 3;                 // TMEM lane quarter this warp may access
        float* tileS = epi_sm + q * 32 * EPI_LD;
        float* biasS = epi_sm + 4 * 32 * EPI_LD + q * EPI_BIAS;
        int acc = 0; uint32_t accphase = 0;
        for (int tile0 = blockIdx.x; tile0 < total_tiles; tile0 += gridDim.x) {
            const int split = tile0 % splitk, tile = tile0 / splitk;
            const int ntile = tile % g.nt, mrem = tile / g.nt;
            const int mtile = mrem % g.mt_per_b, b = mrem / g.mt_per_b + g.b_first;
            if (g.lens && mtile * BM >= g.lens[b]) continue;
            mbar_wait(tfull_bar(acc), accphase);
            tc_fence_after();
            if (EPI == 1) {
                // ---- fused LSTM cell: this thread = batch row, tile = units [ntile*16, +16) x gates (i|f|g|o) ----
                const LasLstmEpi& le = g.le;
                const int H = le.H, bb = mtile * BM + q * 32 + lane, ub = ntile * 16;
                uint32_t v0[32], v1[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN, v0);          // [i(16) | f(16)]
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + 32, v1);     // [g(16) | o(16)]
                if (bb < g.R) {
                    const float* tr = nullptr;
                    if (le.tab) {
                        int tok;
                        if (le.t == 0) tok = le.sos_idx;
                        else if (le.use_gold) tok = le.y[(long long)bb * le.ld_y + le.t - 1];
                        else tok = le.chars_prev[bb];
                        if (le.tok_out && ntile == 0) le.tok_out[bb] = tok;
                        tr = le.tab + (long long)tok * 4 * H;
                    }
                    const float* cp = le.c_prev + (long long)bb * le.ld_cp;
                    float* co = le.c_out + (long long)bb * le.ld_co;
                    float* gout = le.G + (long long)bb * 4 * H;
#pragma unroll
                    for (int ul = 0; ul < 16; ++ul) {
                        const int u = ub + ul;
                        if (u < H) {
                            float p0 = __uint_as_float(v0[ul]), p1 = __uint_as_float(v0[16 + ul]);
                            float p2 = __uint_as_float(v1[ul]), p3 = __uint_as_float(v1[16 + ul]);
                            if (tr) { p0 += tr[u]; p1 += tr[H + u]; p2 += tr[2 * H + u]; p3 += tr[3 * H + u]; }
                            if (le.bias1) { p0 += le.bias1[u]; p1 += le.bias1[H + u]; p2 += le.bias1[2 * H + u]; p3 += le.bias1[3 * H + u]; }
                            if (le.bias2) { p0 += le.bias2[u]; p1 += le.bias2[H + u]; p2 += le.bias2[2 * H + u]; p3 += le.bias2[3 * H + u]; }
                            const float gi = tc_sigmoid_fast(p0), gf = tc_sigmoid_fast(p1), gg = tc_tanh_fast(p2), go = tc_sigmoid_fast(p3);
                            const float c = fmaf(gf, cp[u], gi * gg);
                            float h = go * tc_tanh_fast(c);
                            if (le.mask) h *= le.mask[(long long)bb * H + u];
                            gout[u] = gi; gout[H + u] = gf; gout[2 * H + u] = gg; gout[3 * H + u] = go;
                            co[u] = c;
                            le.h1[(long long)bb * le.ld_h1 + u] = h;
                            if (le.h2) le.h2[(long long)bb * le.ld_h2 + u] = h;
                            if (le.h1b) le.h1b[(long long)bb * le.ld_h1b + u] = __float2bfloat16(h);
                            if (le.h2b) le.h2b[(long long)bb * le.ld_h2b + u] = __float2bfloat16(h);
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty_bar(acc));
                if (++acc == 2) { acc = 0; accphase ^= 1u; }
                continue;
            }
            const int r0 = mtile * BM + q * 32;                     // first row of this warp's 32-row band
            float* cbase;
            long long ldo;
            if (splitk > 1) { cbase = g.Cpart + (long long)split * g.R * g.ldp; ldo = g.ldp; }
            else { cbase = g.C + (long long)b * g.c_bs; ldo = g.ldc; }
            const bool plain = splitk > 1;                          // partials carry no bias / accumulate
            const bool has_bias = !plain && (g.bias1 || g.bias2);
            const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
            uint32_t va[32], vb[32];
            tmem_ld32_issue(tbase, va);                             // chunk 0 is on its way while the bias sums are fetched
            if (has_bias) {
                // bias1 + bias2 of the tile's columns, once per tile (they used to be two dependent global loads per 32-column chunk in
                // front of every drain: 8 % of all warp-stall samples of the K = 64 base-layer projection, which is all epilogue)
                // (every load of a lane is issued before the first add: one round trip per tile -- a scalar loop here was still 11 % of the samples)
                constexpr int NV = BN / 128;                         // 128-bit pieces per lane: 2 (BN = 256) or 0 (BN = 64: scalar below)
                float4 t1[NV > 0 ? NV : 1], t2[NV > 0 ? NV : 1];
                const bool vec = (g.N % 4 == 0) && ((reinterpret_cast<uintptr_t>(g.bias1) | reinterpret_cast<uintptr_t>(g.bias2)) & 15) == 0;
                if (NV > 0 && vec) {
#pragma unroll
                    for (int i = 0; i < NV; ++i) {
                        const int n = ntile * BN + (i * 32 + lane) * 4;
                        const bool in = n < g.N;
                        t1[i] = (in && g.bias1) ? *reinterpret_cast<const float4*>(g.bias1 + n) : make_float4(0.f, 0.f, 0.f, 0.f);
                        t2[i] = (in && g.bias2) ? *reinterpret_cast<const float4*>(g.bias2 + n) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
#pragma unroll
                    for (int i = 0; i < NV; ++i)
                        *reinterpret_cast<float4*>(biasS + (i * 32 + lane) * 4) =
                            make_float4(t1[i].x + t2[i].x, t1[i].y + t2[i].y, t1[i].z + t2[i].z, t1[i].w + t2[i].w);
                } else {
                    for (int j = lane; j < BN; j += 32) {
                        const int n = ntile * BN + j;
                        float v = 0.f;
                        if (n < g.N) { if (g.bias1) v += g.bias1[n]; if (g.bias2) v += g.bias2[n]; }
                        biasS[j] = v;
                    }
                }
                __syncwarp();
            }
            // one 32-column chunk: stage thread `lane`'s row (tileS[lane][0..31]), then drain 4 rows x 128 contiguous bytes per instruction
            auto chunk = [&](const uint32_t (&v)[32], int c) {
                const int n0 = ntile * BN + c * 32;
                if (n0 >= g.N) return;                              // warp-uniform
                // accumulate form: C += tile as fire-and-forget 128-bit reductions at L2 (red.global.add.v4.f32; one writer per element per
                // launch and launches of one stream are ordered, so the sum is the same single fp32 add as load + add + store, and
                // deterministic).  Loads inside the drain loop are one dependent L2 round trip per store -- no load may move above the
                // preceding store to the same array -- and made a K = 1024 tile's epilogue three times as long as its UMMAs (measured
                // with the direction-half gate tiles: 37.8 -> 42.1 ms per train step); eight prefetched loads per 32-column block still
                // left it epilogue-bound (13 us per output tile against 8 us of UMMAs).
                const bool acc_full = !plain && g.accumulate && n0 + 32 <= g.N;
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4*>(tileS + lane * EPI_LD + j) =
                        make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                __syncwarp();
                const int cq = (lane & 7) * 4, rsub = lane >> 3;
                const int n = n0 + cq;
                if (n0 + 32 <= g.N) {
                    float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (has_bias) bb = *reinterpret_cast<const float4*>(biasS + c * 32 + cq);
                    float4 o[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) o[i] = *reinterpret_cast<const float4*>(tileS + (i * 4 + rsub) * EPI_LD + cq);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int rr = i * 4 + rsub;
                        if (r0 + rr < g.R) {
                            o[i].x += bb.x; o[i].y += bb.y; o[i].z += bb.z; o[i].w += bb.w;
                            float4* dst = reinterpret_cast<float4*>(cbase + (long long)(r0 + rr) * ldo + n);
                            if (acc_full)
                                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(o[i].x), "f"(o[i].y), "f"(o[i].z), "f"(o[i].w) : "memory");
                            else
                                *dst = o[i];
                        }
                    }
                } else {
                    for (int i = 0; i < 8; ++i) {
                        const int rr = i * 4 + rsub;
                        if (r0 + rr >= g.R) continue;
                        for (int e = 0; e < 4 && n + e < g.N; ++e) {
                            float o = tileS[rr * EPI_LD + cq + e];
                            if (has_bias) o += biasS[c * 32 + cq + e];
                            float* dst = cbase + (long long)(r0 + rr) * ldo + n + e;
                            if (!plain && g.accumulate) o += *dst;
                            *dst = o;
                        }
                    }
                }
                __syncwarp();
            };
            // the accumulator read of chunk c + 1 is in flight while chunk c is staged and drained
#pragma unroll 1
            for (int c = 0; c < BN / 32; c += 2) {
                tmem_ld_wait();
                tmem_ld32_issue(tbase + (c + 1) * 32, vb);          // BN / 32 is even (2 or 8)
                chunk(va, c);
                tmem_ld_wait();
                if (c + 2 < BN / 32) tmem_ld32_issue(tbase + (c + 2) * 32, va);
                chunk(vb, c + 1);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
            if (++acc == 2) { acc = 0; accphase ^= 1u; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ---- host side -----------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = (EncodeTiledFn)p;
    return fn;
}

// 3-D bf16 tensor: dim0 contiguous (d0 elements), dim1 rows (d1, stride s1 elements), dim2 batches (d2, stride s2 elements)
int make_map(CUtensorMap* m, const void* ptr, long long d0, long long d1, long long d2, long long s1, long long s2, int box0, int box1) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { las_set_error("cuTensorMapEncodeTiled entry point not available"); return LAS_ERR_CUDA; }
    LAS_CHECK_ARG(((uintptr_t)ptr & 15) == 0, "gemm_tc: operand pointer %p not 16-byte aligned", ptr);
    LAS_CHECK_ARG((s1 * 2) % 16 == 0 && (d2 <= 1 || (s2 * 2) % 16 == 0), "gemm_tc: operand strides (%lld, %lld elements) must be multiples of 8",
                  s1, s2);
    cuuint64_t dims[3] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)(d2 < 1 ? 1 : d2)};
    cuuint64_t strides[2] = {(cuuint64_t)s1 * 2, (cuuint64_t)((d2 <= 1 ? s1 * d1 : s2) * 2)};
    cuuint32_t box[3] = {(cuuint32_t)box0, (cuuint32_t)box1, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        las_set_error("cuTensorMapEncodeTiled failed (%d) dims=(%lld,%lld,%lld) strides=(%lld,%lld) box=(%d,%d)", (int)r, d0, d1, d2, s1, s2,
                      box0, box1);
        return LAS_ERR_CUDA;
    }
    return LAS_OK;
}

// C[m][n] (+)= sum_s part[s][m][n] (+ biases): deterministic split-K reduction
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ part, long long ldp, int splitk, int R, int N, float* __restrict__ C,
                                                            long long ldc, const float* __restrict__ bias1, const float* __restrict__ bias2,
                                                            int accumulate) {
    const long long total = (long long)R * (N / 4);
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int m = (int)(i / (N / 4)), n = (int)(i % (N / 4)) * 4;
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int sidx = 0; sidx < splitk; ++sidx) {
            const float4 t = *reinterpret_cast<const float4*>(part + ((long long)sidx * R + m) * ldp + n);
            o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w;
        }
        if (bias1) { o.x += bias1[n]; o.y += bias1[n + 1]; o.z += bias1[n + 2]; o.w += bias1[n + 3]; }
        if (bias2) { o.x += bias2[n]; o.y += bias2[n + 1]; o.z += bias2[n + 2]; o.w += bias2[n + 3]; }
        float4* dst = reinterpret_cast<float4*>(C + (long long)m * ldc + n);
        if (accumulate) { const float4 old = *dst; o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w; }
        *dst = o;
    }
}

// LasGemmTc::max_ctas of the call being launched (0 = one CTA per SM): lets a GEMM that runs beside a persistent recurrence kernel on
// another stream take only the SMs that kernel leaves free
thread_local int t_max_ctas = 0;

template <bool A_MN, bool B_MN, int BN, int EPI = 0>
int launch_tc(const CUtensorMap& ta, const CUtensorMap& tb, const TcArgs& g, cudaStream_t st) {
    auto kern = gemm_bf16_tc_kernel<A_MN, B_MN, BN, EPI>;
    constexpr int SMEM_BYTES = Cfg<BN>::SMEM_BYTES;
    static bool attr_set = false;
    if (!attr_set) {
        LAS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        attr_set = true;
    }
    const int total = g.NB * g.mt_per_b * g.nt * (g.splitk > 1 ? g.splitk : 1);
    int grid = las_device_info()->num_sms;
    if (t_max_ctas > 0 && grid > t_max_ctas) grid = t_max_ctas;
    if (grid > total) grid = total;
    LAS_CUDA(las_launch(kern, dim3(grid), dim3(NTHREADS), SMEM_BYTES, st, ta, tb, g));
    LAS_LAUNCH_CHECK();
    return LAS_OK;
}

}  // namespace

// ---- prepared plans (internal C++ API used by the decoder loop): tensor maps are encoded once, every step only launches
struct LasTcPlan {
    CUtensorMap ta, tb;
    TcArgs g;
    int variant;     // 0: A,B K-major (C = A.B^T) ; 1: B MN-major (C = A.B)
    int bn;          // N tile: 64, or 32 for the K-major form (twice the CTAs, each ingesting a smaller weight slice)
};
size_t las_tc_plan_bytes() { return sizeof(LasTcPlan); }

// A: (K contiguous, M rows [a_s1], a_batches [a_s2]); one batch is selected per launch.  Always the BN = 64 tiling.
int las_tc_plan_make(void* plan_mem, const void* A, const void* B, int M, int N, int K, int a_batches, long long a_s1, long long a_s2,
                     long long b_s1, int b_mn_major) {
    LasTcPlan* p = new (plan_mem) LasTcPlan();
    int rc = make_map(&p->ta, A, K, M, a_batches, a_s1, a_s2, BK, BM);
    if (rc) return rc;
    // these GEMMs have one M tile (M = batch <= 128): every CTA pulls the whole activation matrix plus its own weight slice.
    // N tiles of 32 (LAS_DEC_BN=32) put twice as many SMs to work on half the weight slice each; measured no faster than 64
    // (decoder forward loop 10.47 vs 10.30 ms: the shared activation tile dominates what each SM ingests), so 64 is the default.
    const char* e = getenv("LAS_DEC_BN");
    const char* fe = getenv("LAS_DEC_FUSE");          // the opt-in fused LSTM epilogue is written for 64-column tiles
    const int want = (fe && atoi(fe) == 1) ? 64 : (e ? atoi(e) : 64);
    p->bn = (!b_mn_major && want == 32 && N % 32 == 0 && ceil_div(N, 32) <= las_device_info()->num_sms) ? 32 : 64;
    if (!b_mn_major) rc = make_map(&p->tb, B, K, N, 1, b_s1, 0, BK, p->bn);
    else rc = make_map(&p->tb, B, N, K, 1, b_s1, 0, 64, 64);
    if (rc) return rc;
    p->variant = b_mn_major ? 1 : 0;
    TcArgs& g = p->g;
    g = TcArgs{};
    g.R = M; g.NB = 1; g.N = N; g.mt_per_b = ceil_div(M, BM); g.nt = ceil_div(N, p->bn); g.kt_per_b = ceil_div(K, BK); g.KB = 1;
    return LAS_OK;
}

int las_tc_plan_launch(const void* plan_mem, int a_batch, float* C, long long ldc, const float* bias1, const float* bias2, void* stream) {
    const LasTcPlan* p = (const LasTcPlan*)plan_mem;
    TcArgs g = p->g;
    g.C = C; g.ldc = ldc; g.c_bs = 0; g.bias1 = bias1; g.bias2 = bias2; g.b_first = a_batch;
    LAS_CHECK_ARG(ldc % 4 == 0 && ((uintptr_t)C & 15) == 0, "tc plan: C must be 16-byte aligned with ldc %% 4 == 0");
    LasProfScope prof(LAS_PROF_GEMM_OTHER, stream, 2.0 * g.R * (double)g.N * g.kt_per_b * BK);
    // the c_bs * b term must vanish for the selected batch: C is already the step's output
    g.c_bs = 0;
    if (p->variant == 0) {
        if (p->bn == 32) return launch_tc<false, false, 32>(p->ta, p->tb, g, (cudaStream_t)stream);
        return launch_tc<false, false, 64>(p->ta, p->tb, g, (cudaStream_t)stream);
    }
    return launch_tc<false, true, 64>(p->ta, p->tb, g, (cudaStream_t)stream);
}

// plan launch with the K range split over `splitk` CTAs per output tile; the partial matrices (splitk, M, ldp) are left
// un-reduced in Cpart: the decoder's pointwise consumers (cell kernels, attention backward) add them up as they read
int las_tc_plan_launch_split(const void* plan_mem, int a_batch, float* Cpart, long long ldp, int splitk, void* stream) {
    const LasTcPlan* p = (const LasTcPlan*)plan_mem;
    TcArgs g = p->g;
    LAS_CHECK_ARG(splitk >= 1 && splitk <= g.kt_per_b && ldp % 4 == 0 && ((uintptr_t)Cpart & 15) == 0, "tc plan split: bad split / ldp");
    g.b_first = a_batch;
    if (splitk == 1) { g.C = Cpart; g.ldc = ldp; g.c_bs = 0; }
    else { g.splitk = splitk; g.Cpart = Cpart; g.ldp = ldp; }
    LasProfScope prof(LAS_PROF_GEMM_OTHER, stream, 2.0 * g.R * (double)g.N * g.kt_per_b * BK);
    if (p->variant == 0) {
        if (p->bn == 32) return launch_tc<false, false, 32>(p->ta, p->tb, g, (cudaStream_t)stream);
        return launch_tc<false, false, 64>(p->ta, p->tb, g, (cudaStream_t)stream);
    }
    return launch_tc<false, true, 64>(p->ta, p->tb, g, (cudaStream_t)stream);
}
int las_tc_plan_tiles(const void* plan_mem) { const LasTcPlan* p = (const LasTcPlan*)plan_mem; return p->g.mt_per_b * p->g.nt; }
int las_tc_plan_kiters(const void* plan_mem) { return ((const LasTcPlan*)plan_mem)->g.kt_per_b; }

// plan launch with the fused LSTM-cell epilogue (plan must be the K-major-B form with permuted weight rows, N = 4H)
int las_tc_plan_launch_lstm(const void* plan_mem, int a_batch, const LasLstmEpi* le, void* stream) {
    const LasTcPlan* p = (const LasTcPlan*)plan_mem;
    LAS_CHECK_ARG(p->variant == 0 && p->bn == 64 && le && le->H % 16 == 0 && p->g.N == 4 * le->H, "tc plan (lstm epilogue): bad plan / H");
    TcArgs g = p->g;
    g.C = nullptr; g.b_first = a_batch; g.le = *le;
    LasProfScope prof(LAS_PROF_GEMM_OTHER, stream, 2.0 * g.R * (double)g.N * g.kt_per_b * BK);
    return launch_tc<false, false, 64, 1>(p->ta, p->tb, g, (cudaStream_t)stream);
}

// dst[n'][k] (bf16) = src[perm(n')][k]: rows regrouped so each 64-row block is 16 units x (i|f|g|o); src is gate-major (4H, K)
__global__ void __launch_bounds__(256) permute_cast_kernel(const float* __restrict__ src, long long ld_src, __nv_bfloat16* __restrict__ dst,
                                                           int H, int K) {
    const long long total = (long long)4 * H * K;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int np = (int)(i / K), k = (int)(i - (long long)np * K);
        const int nt = np >> 6, j = np & 63, gidx = j >> 4, ul = j & 15;
        dst[i] = __float2bfloat16(src[(long long)(gidx * H + nt * 16 + ul) * ld_src + k]);
    }
}
int las_permute_cast_lstm_rows(const float* src, long long ld_src, void* dst, int H, int K, void* stream) {
    LAS_CHECK_ARG(src && dst && H % 16 == 0 && K >= 1, "permute_cast: bad arguments");
    const long long total = (long long)4 * H * K;
    int grid = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
    permute_cast_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, ld_src, (__nv_bfloat16*)dst, H, K);
    LAS_LAUNCH_CHECK();
    return LAS_OK;
}

extern "C" int las_gemm_bf16_tc(const LasGemmTc* d, void* stream) {
    LAS_CHECK_ARG(d != nullptr && d->A && d->B && d->C, "gemm_tc: null descriptor / operand");
    LAS_CHECK_ARG(d->M >= 1 && d->N >= 1 && d->K >= 1, "gemm_tc: bad dims M=%d N=%d K=%d", d->M, d->N, d->K);
    LAS_CHECK_ARG(d->a_batches >= 1 && d->k_batches >= 1, "gemm_tc: bad batch counts");
    LAS_CHECK_ARG(d->ldc % 4 == 0 && ((uintptr_t)d->C & 15) == 0 && d->c_bs % 4 == 0, "gemm_tc: C must be 16-byte aligned with ldc %% 4 == 0");
    int rc = las_set_device_of(d->C);
    if (rc) return rc;
    struct CapScope { CapScope(int v) { t_max_ctas = v; } ~CapScope() { t_max_ctas = 0; } } cap_scope(d->max_ctas);
    cudaStream_t st = (cudaStream_t)stream;
    CUtensorMap ta, tb;
    TcArgs g{};
    g.C = d->C; g.bias1 = d->bias1; g.bias2 = d->bias2; g.c_bs = d->c_bs; g.ldc = d->ldc;
    g.N = d->N; g.accumulate = d->accumulate; g.lens = d->lens;
    // measured on B200: a kind::f16 UMMA with one fp16 and one bf16 operand raises an illegal-instruction error, so both or neither
    LAS_CHECK_ARG((d->a_f16 != 0) == (d->b_f16 != 0), "gemm_tc: both operands must have the same 16-bit format (fp16 x bf16 is not a legal tcgen05 kind::f16 pair)");
    g.fmt_clear = (d->a_f16 ? (1u << 7) : 0u) | (d->b_f16 ? (1u << 10) : 0u);
    const double flops = d->prof_flops > 0 ? d->prof_flops : 2.0 * d->M * (double)d->N * d->K * d->a_batches * d->k_batches;
    LasProfScope prof(d->prof_tag == 2 ? LAS_PROF_GEMM_GATES_SIDE
                                       : d->prof_tag == 1 ? (d->max_ctas > 0 ? LAS_PROF_GEMM_GATES_SIDE : LAS_PROF_GEMM_GATES) : LAS_PROF_GEMM_OTHER,
                      stream, flops);
    // narrow N tiles when the 128x256 tiling would leave most SMs idle (decoder-step GEMMs: M = batch)
    const long long tiles256 = (long long)ceil_div(d->M, BM) * ceil_div(d->N, 256) * d->a_batches;
    const bool narrow = tiles256 < 40 && !(d->a_mn_major && d->splitk > 1);
    const int BNsel = narrow ? 64 : 256;
    g.nt = ceil_div(d->N, BNsel);
    if (!d->a_mn_major) {
        // A: (K contiguous, M rows [stride a_s1], a_batches [stride a_s2]); reduction is a single K range
        LAS_CHECK_ARG(d->k_batches == 1, "gemm_tc: K-major A cannot have K batches");
        long long kspan = d->K;         // extent of the contiguous dimension the tensor maps cover
        if (d->k_chunk > 0) {
            // chunked reduction: K = nchunk * k_chunk elements, chunk c starts at c * k_chunk_stride in A's and in B's K dimension
            LAS_CHECK_ARG(!d->b_mn_major && d->k_chunk % BK == 0 && d->K % d->k_chunk == 0 && d->k_chunk_stride >= d->k_chunk &&
                              d->k_chunk_stride % 8 == 0,
                          "gemm_tc: chunked K needs K-major A and B, k_chunk %% %d == 0, K %% k_chunk == 0, stride >= chunk", BK);
            g.kc_iters = d->k_chunk / BK; g.kc_stride = d->k_chunk_stride;
            kspan = (long long)(d->K / d->k_chunk - 1) * d->k_chunk_stride + d->k_chunk;
        }
        rc = make_map(&ta, d->A, kspan, d->M, d->a_batches, d->a_s1, d->a_s2, BK, BM);
        if (rc) return rc;
        g.R = d->M; g.NB = d->a_batches; g.mt_per_b = ceil_div(d->M, BM); g.kt_per_b = ceil_div(d->K, BK); g.KB = 1;
        if (!d->b_mn_major) {
            rc = make_map(&tb, d->B, kspan, d->N, 1, d->b_s1, 0, BK, BNsel);        // B: (K contiguous, N rows)
            if (rc) return rc;
            return narrow ? launch_tc<false, false, 64>(ta, tb, g, st) : launch_tc<false, false, 256>(ta, tb, g, st);
        }
        rc = make_map(&tb, d->B, d->N, d->K, 1, d->b_s1, 0, 64, 64);               // B: (N contiguous, K rows)
        if (rc) return rc;
        return narrow ? launch_tc<false, true, 64>(ta, tb, g, st) : launch_tc<false, true, 256>(ta, tb, g, st);
    }
    // A: (M contiguous, K rows [stride a_s1], k_batches [stride a_s2]); B: (N contiguous, K rows [b_s1], k_batches [b_s2])
    LAS_CHECK_ARG(d->b_mn_major && d->a_batches == 1, "gemm_tc: MN-major A needs MN-major B and a single M batch");
    rc = make_map(&ta, d->A, d->M, d->K, d->k_batches, d->a_s1, d->a_s2, 64, 64);
    if (rc) return rc;
    rc = make_map(&tb, d->B, d->N, d->K, d->k_batches, d->b_s1, d->b_s2, 64, 64);
    if (rc) return rc;
    g.R = d->M; g.NB = 1; g.mt_per_b = ceil_div(d->M, BM); g.kt_per_b = ceil_div(d->K, BK); g.KB = d->k_batches;
    g.lens = nullptr;
    if (d->splitk > 1) {
        LAS_CHECK_ARG(d->workspace != nullptr && d->N % 4 == 0, "gemm_tc: split-K needs a workspace and N %% 4 == 0");
        g.splitk = d->splitk; g.Cpart = d->workspace; g.ldp = (d->N + 3) & ~3;
        rc = launch_tc<true, true, 256>(ta, tb, g, st);
        if (rc) return rc;
        const long long total = (long long)d->M * (d->N / 4);
        int grid = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
        splitk_reduce_kernel<<<grid, 256, 0, st>>>(g.Cpart, g.ldp, g.splitk, d->M, d->N, d->C, d->ldc, d->bias1, d->bias2, d->accumulate);
        LAS_LAUNCH_CHECK();
        return LAS_OK;
    }
    return narrow ? launch_tc<true, true, 64>(ta, tb, g, st) : launch_tc<true, true, 256>(ta, tb, g, st);
}

// ---- fp32 -> bf16 cast with optional column padding: dst[r][c] = c < cols ? src[r*ld_src + c] : 0 ---------------------
template <bool F16>
__global__ void __launch_bounds__(256) cast_bf16_kernel(const float* __restrict__ src, long long ld_src, long long inner, long long bs,
                                                        __nv_bfloat16* __restrict__ dst, long long ld_dst, long long rows, int cols,
                                                        int cols_pad) {
    const long long total = rows * (long long)(cols_pad / 2);
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const long long r = i / (cols_pad / 2);
        const int c = (int)(i - r * (cols_pad / 2)) * 2;
        const float* srow = src + (inner > 0 ? (r / inner) * bs + (r % inner) * ld_src : r * ld_src);
        const float a = c < cols ? srow[c] : 0.f;
        const float b = c + 1 < cols ? srow[c + 1] : 0.f;
        if (F16) *reinterpret_cast<__half2*>(dst + r * ld_dst + c) = __floats2half2_rn(a, b);
        else *reinterpret_cast<__nv_bfloat162*>(dst + r * ld_dst + c) = __floats2bfloat162_rn(a, b);
    }
}

extern "C" int las_cast_f32_to_bf16(const float* src, long long ld_src, long long inner, long long bs, void* dst, long long ld_dst,
                                    long long rows, int cols, int cols_pad, void* stream) {
    LAS_CHECK_ARG(src && dst && rows >= 0 && cols >= 1 && cols_pad >= cols && cols_pad % 2 == 0 && ld_dst % 2 == 0,
                  "cast_bf16: bad arguments");
    if (rows == 0) return LAS_OK;
    int rc = las_set_device_of(dst);
    if (rc) return rc;
    long long total = rows * (long long)(cols_pad / 2);
    int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    cast_bf16_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(src, ld_src, inner, bs, (__nv_bfloat16*)dst, ld_dst, rows, cols, cols_pad);
    LAS_LAUNCH_CHECK();
    return LAS_OK;
}

// same, IEEE half destination (decoder forward operands: |h| <= 1/(1-p), weights O(1) -- 10 mantissa bits instead of 7)
extern "C" int las_cast_f32_to_f16(const float* src, long long ld_src, long long inner, long long bs, void* dst, long long ld_dst,
                                   long long rows, int cols, int cols_pad, void* stream) {
    LAS_CHECK_ARG(src && dst && rows >= 0 && cols >= 1 && cols_pad >= cols && cols_pad % 2 == 0 && ld_dst % 2 == 0,
                  "cast_f16: bad arguments");
    if (rows == 0) return LAS_OK;
    int rc = las_set_device_of(dst);
    if (rc) return rc;
    long long total = rows * (long long)(cols_pad / 2);
    int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    cast_bf16_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(src, ld_src, inner, bs, (__nv_bfloat16*)dst, ld_dst, rows, cols, cols_pad);
    LAS_LAUNCH_CHECK();
    return LAS_OK;
}
