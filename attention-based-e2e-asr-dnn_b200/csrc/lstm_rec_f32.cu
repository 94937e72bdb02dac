// Persistent LSTM recurrence, fp32 parity mode (forward + BPTT).
//
// Replaces the time loop inside nn.LSTM on a PackedSequence (reference src/modules.py:78-82 and :187-191; gate
// equations torch nn/modules/rnn.py:842-847).  One cooperative launch runs all T timesteps of BOTH directions:
//   * grid = (NC, ndir); CTA (c, dir) owns hidden units [c*U, (c+1)*U) of direction dir and keeps the matching
//     4U rows of W_hh resident in shared memory for the whole sequence;
//   * per timestep the CTA computes gates = xgates[b,t] + W_hh_slice . h_{t-1}[b,:], applies the gate
//     nonlinearities and the cell update in registers, writes h_t/c_t, and the CTAs of one direction meet at a
//     per-direction barrier (release/acquire counter in global memory) before the next timestep;
//   * PackedSequence semantics without packing: a row takes part in step t only while t < len[b]; the reverse
//     direction walks t = T-1..0 so each row starts at its own last valid frame with zero state; positions
//     t >= len[b] are written as exact zeros (pad_packed_sequence);
//   * the locked-dropout mask (B, ndir*H) of src/modules.py:61-64 is applied on the output write only -- the
//     recurrent state stays un-dropped, as in the reference where dropout follows the whole layer.
// State lives in the saved tensors themselves: hs_pad / cs_pad are (B, T+2, ndir*H) with frame t+1 holding time t and
// frames 0 and T+1 zero, so "previous" is frame t (forward) or t+2 (reverse) with no special cases -- the same
// padded frames give the shifted h_{t-1} operand of the dW_hh GEMM in backward.
#include "las_common.cuh"
#include "las_b200.h"
#include <cooperative_groups.h>

namespace {

constexpr int U = 8;          // hidden units per CTA
constexpr int NTHREADS = 256;
constexpr int NB = NTHREADS / U;   // 32 batch lanes
constexpr int RB = 3;              // batch rows per thread per chunk
constexpr int CHUNK = NB * RB;     // 96 rows
constexpr int PAD = 4;

struct RecArgs {
    float* gates;          // (B, T, ndir, 4H): in x-gates (+biases); out activated gates / in bwd: out dgates
    const float* w_hh;     // (ndir, 4H, H)
    const int* lens;       // (B)
    const float* mask;     // (B, ndir*H) or null
    float* out;            // (B, T, ndir*H) or null
    const float* dout;     // bwd: (B, T, ndir*H) contiguous
    float* hs_pad;         // (B, T+2, ndir*H)
    float* cs_pad;         // (B, T+2, ndir*H)
    float* dstate;         // bwd workspace: dh (ndir,B,H) then dc (ndir,B,H)
    unsigned* ctr;         // (ndir) zeroed
    int B, T, H, ndir, NC, KC;
};

__device__ __forceinline__ void dir_barrier(unsigned* ctr, unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        red_release_gpu_add(ctr, 1u);
        while (ld_acquire_gpu(ctr) < target) { __nanosleep(20); }
    }
    __syncthreads();
}

// stage rows [b0, b0+CHUNK) x cols [k0, k0+KC) of a (row-strided) matrix into smem[CHUNK][KC+PAD] with cp.async
__device__ __forceinline__ void stage_rows(float* dst, const float* src, long long row_stride, int b0, int B, int KC) {
    const int pieces = KC / 4;
    const int total = CHUNK * pieces;
    for (int idx = threadIdx.x; idx < total; idx += NTHREADS) {
        int r = idx / pieces, p = idx - r * pieces;
        bool valid = (b0 + r) < B;
        const float* s = src + (valid ? (long long)(b0 + r) * row_stride + p * 4 : 0);
        cp_async16_zfill(dst + r * (KC + PAD) + p * 4, s, valid);
    }
}

__global__ void __launch_bounds__(NTHREADS, 1) lstm_rec_fwd_f32_kernel(RecArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int H = a.H, T = a.T, B = a.B, KC = a.KC;
    const int F = a.ndir * H;
    float* Wsl = smem;                                 // [4U][H+PAD]
    float* hst = smem + 4 * U * (H + PAD);             // [2][CHUNK][KC+PAD]
    const int dir = blockIdx.y, u0 = blockIdx.x * U;
    const int tid = threadIdx.x;
    const int ul = tid % U, bl = tid / U;
    const int u = u0 + ul;
    const long long frame_stride = F;
    const long long brow = (long long)(T + 2) * F;

    // W_hh slice -> smem (row g*U+ul <- W_hh[dir][g*H + u0 + ul][:])
    for (int idx = tid; idx < 4 * U * (H / 4); idx += NTHREADS) {
        int row = idx / (H / 4), k4 = idx - row * (H / 4);
        int g = row / U, uu = row - g * U;
        const float4 v = *reinterpret_cast<const float4*>(a.w_hh + ((long long)dir * 4 * H + g * H + u0 + uu) * H + k4 * 4);
        *reinterpret_cast<float4*>(Wsl + row * (H + PAD) + k4 * 4) = v;
    }
    // zero the two pad frames of this CTA's (dir, unit) slice
    for (int idx = tid; idx < B * U * 2; idx += NTHREADS) {
        int b = idx / (2 * U), r = idx - b * 2 * U;
        int uu = r % U, which = r / U;
        long long off = (long long)b * brow + (which ? (long long)(T + 1) * F : 0) + dir * H + u0 + uu;
        a.hs_pad[off] = 0.f;
        a.cs_pad[off] = 0.f;
    }
    __syncthreads();

    const int nsub = H / KC;
    for (int s = 0; s < T; ++s) {
        const int t = (dir == 0) ? s : (T - 1 - s);
        const int fprev = (dir == 0) ? t : t + 2;
        const int fcur = t + 1;
        for (int b0 = 0; b0 < B; b0 += CHUNK) {
            float acc[RB][4];
#pragma unroll
            for (int j = 0; j < RB; ++j)
#pragma unroll
                for (int g = 0; g < 4; ++g) acc[j][g] = 0.f;
            if (s > 0) {
                const float* hsrc = a.hs_pad + (long long)fprev * frame_stride + dir * H;
                stage_rows(hst, hsrc, brow, b0, B, KC);
                cp_async_commit();
                for (int sub = 0; sub < nsub; ++sub) {
                    float* cur = hst + (sub & 1) * CHUNK * (KC + PAD);
                    if (sub + 1 < nsub) {
                        stage_rows(hst + ((sub + 1) & 1) * CHUNK * (KC + PAD), hsrc + (sub + 1) * KC, brow, b0, B, KC);
                        cp_async_commit();
                        cp_async_wait<1>();
                    } else {
                        cp_async_wait<0>();
                    }
                    __syncthreads();
                    const float* wbase = Wsl + ul * (H + PAD) + sub * KC;
                    for (int k = 0; k < KC; k += 4) {
                        float4 w4[4], h4[RB];
#pragma unroll
                        for (int g = 0; g < 4; ++g)
                            w4[g] = *reinterpret_cast<const float4*>(wbase + g * U * (H + PAD) + k);
#pragma unroll
                        for (int j = 0; j < RB; ++j)
                            h4[j] = *reinterpret_cast<const float4*>(cur + (bl + NB * j) * (KC + PAD) + k);
#pragma unroll
                        for (int j = 0; j < RB; ++j)
#pragma unroll
                            for (int g = 0; g < 4; ++g) {
                                acc[j][g] = fmaf(w4[g].x, h4[j].x, acc[j][g]);
                                acc[j][g] = fmaf(w4[g].y, h4[j].y, acc[j][g]);
                                acc[j][g] = fmaf(w4[g].z, h4[j].z, acc[j][g]);
                                acc[j][g] = fmaf(w4[g].w, h4[j].w, acc[j][g]);
                            }
                    }
                    __syncthreads();
                }
            }
            // gate nonlinearities + cell update (fused epilogue)
#pragma unroll
            for (int j = 0; j < RB; ++j) {
                const int b = b0 + bl + NB * j;
                if (b >= B) continue;
                const long long so = (long long)b * brow + (long long)fcur * F + dir * H + u;
                const bool valid = t < a.lens[b];
                float h = 0.f, c = 0.f;
                if (valid) {
                    float* gp = a.gates + (((long long)b * T + t) * a.ndir + dir) * 4 * H + u;
                    const float gi = sigmoidf_acc(acc[j][0] + gp[0]);
                    const float gf = sigmoidf_acc(acc[j][1] + gp[H]);
                    const float gg = tanhf(acc[j][2] + gp[2 * H]);
                    const float go = sigmoidf_acc(acc[j][3] + gp[3 * H]);
                    const float cprev = (s == 0) ? 0.f : a.cs_pad[(long long)b * brow + (long long)fprev * F + dir * H + u];
                    c = fmaf(gf, cprev, gi * gg);
                    h = go * tanhf(c);
                    gp[0] = gi; gp[H] = gf; gp[2 * H] = gg; gp[3 * H] = go;
                }
                a.hs_pad[so] = h;
                a.cs_pad[so] = c;
                if (a.out) {
                    const float m = a.mask ? a.mask[(long long)b * F + dir * H + u] : 1.f;
                    a.out[((long long)b * T + t) * F + dir * H + u] = h * m;
                }
            }
        }
        if (s + 1 < T) dir_barrier(a.ctr + dir, (unsigned)a.NC * (unsigned)(s + 1));
    }
}

// BPTT.  Phase A (pointwise, per (b,u)): dgates for time t from dh = dout*mask + dh_rec and the carried dc.
// Barrier.  Phase B: dh_rec[b,u] = sum_r dgates[b,t,r] * W_hh[r,u] with the column slice W_hh[:, u-slice] resident
// in shared memory.  dgates overwrite the activated gates in place and are what the dW_ih / dW_hh / dX GEMMs consume.
__global__ void __launch_bounds__(NTHREADS, 1) lstm_rec_bwd_f32_kernel(RecArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int H = a.H, T = a.T, B = a.B, KC = a.KC;
    const int F = a.ndir * H, G4 = 4 * H;
    float* WT = smem;                                  // [U][4H+PAD]
    float* gst = smem + U * (G4 + PAD);                // [2][CHUNK][KC+PAD]
    const int dir = blockIdx.y, u0 = blockIdx.x * U;
    const int tid = threadIdx.x;
    const int ul = tid % U, bl = tid / U;
    const int u = u0 + ul;
    const long long brow = (long long)(T + 2) * F;
    float* dh_state = a.dstate + (long long)dir * B * H;
    float* dc_state = a.dstate + (long long)a.ndir * B * H + (long long)dir * B * H;

    for (int idx = tid; idx < U * G4; idx += NTHREADS) {
        int r = idx / U, uu = idx - r * U;
        WT[uu * (G4 + PAD) + r] = a.w_hh[((long long)dir * G4 + r) * H + u0 + uu];
    }
    __syncthreads();

    const int nsub = G4 / KC;
    for (int s = 0; s < T; ++s) {
        const int t = (dir == 0) ? (T - 1 - s) : s;
        const int fprev = (dir == 0) ? t : t + 2;
        const int fcur = t + 1;
        // ---- phase A ----
        for (int b0 = 0; b0 < B; b0 += CHUNK) {
#pragma unroll
            for (int j = 0; j < RB; ++j) {
                const int b = b0 + bl + NB * j;
                if (b >= B) continue;
                float* gp = a.gates + (((long long)b * T + t) * a.ndir + dir) * G4 + u;
                const bool valid = t < a.lens[b];
                float dcn = 0.f;
                if (valid) {
                    const float m = a.mask ? a.mask[(long long)b * F + dir * H + u] : 1.f;
                    float dh = a.dout[((long long)b * T + t) * F + dir * H + u] * m;
                    float dc = 0.f;
                    if (s > 0) {
                        dh += dh_state[(long long)b * H + u];
                        dc = dc_state[(long long)b * H + u];
                    }
                    const float gi = gp[0], gf = gp[H], gg = gp[2 * H], go = gp[3 * H];
                    const float c = a.cs_pad[(long long)b * brow + (long long)fcur * F + dir * H + u];
                    const float cprev = a.cs_pad[(long long)b * brow + (long long)fprev * F + dir * H + u];
                    const float tc = tanhf(c);
                    const float d_o = dh * tc;
                    const float dct = fmaf(dh * go, 1.f - tc * tc, dc);
                    gp[0] = dct * gg * gi * (1.f - gi);
                    gp[H] = dct * cprev * gf * (1.f - gf);
                    gp[2 * H] = dct * gi * (1.f - gg * gg);
                    gp[3 * H] = d_o * go * (1.f - go);
                    dcn = dct * gf;
                } else {
                    gp[0] = 0.f; gp[H] = 0.f; gp[2 * H] = 0.f; gp[3 * H] = 0.f;
                }
                dc_state[(long long)b * H + u] = dcn;
            }
        }
        if (s + 1 == T) break;
        dir_barrier(a.ctr + dir, (unsigned)a.NC * (unsigned)(s + 1));
        // ---- phase B ----
        for (int b0 = 0; b0 < B; b0 += CHUNK) {
            float acc[RB];
#pragma unroll
            for (int j = 0; j < RB; ++j) acc[j] = 0.f;
            // dgates rows of time t: address (b*T + t)*ndir*4H + dir*4H
            const float* gsrc = a.gates + ((long long)t * a.ndir + dir) * G4;
            const long long grow = (long long)T * a.ndir * G4;
            stage_rows(gst, gsrc, grow, b0, B, KC);
            cp_async_commit();
            for (int sub = 0; sub < nsub; ++sub) {
                float* cur = gst + (sub & 1) * CHUNK * (KC + PAD);
                if (sub + 1 < nsub) {
                    stage_rows(gst + ((sub + 1) & 1) * CHUNK * (KC + PAD), gsrc + (sub + 1) * KC, grow, b0, B, KC);
                    cp_async_commit();
                    cp_async_wait<1>();
                } else {
                    cp_async_wait<0>();
                }
                __syncthreads();
                const float* wbase = WT + ul * (G4 + PAD) + sub * KC;
                for (int k = 0; k < KC; k += 4) {
                    const float4 w4 = *reinterpret_cast<const float4*>(wbase + k);
#pragma unroll
                    for (int j = 0; j < RB; ++j) {
                        const float4 g4 = *reinterpret_cast<const float4*>(cur + (bl + NB * j) * (KC + PAD) + k);
                        acc[j] = fmaf(w4.x, g4.x, acc[j]);
                        acc[j] = fmaf(w4.y, g4.y, acc[j]);
                        acc[j] = fmaf(w4.z, g4.z, acc[j]);
                        acc[j] = fmaf(w4.w, g4.w, acc[j]);
                    }
                }
                __syncthreads();
            }
#pragma unroll
            for (int j = 0; j < RB; ++j) {
                const int b = b0 + bl + NB * j;
                if (b < B) dh_state[(long long)b * H + u] = acc[j];
            }
        }
    }
}

int pick_kc(int H) { return (H % 128 == 0) ? 128 : ((H % 64 == 0) ? 64 : 32); }

size_t fwd_smem(int H, int KC) { return sizeof(float) * ((size_t)4 * U * (H + PAD) + (size_t)2 * CHUNK * (KC + PAD)); }
size_t bwd_smem(int H, int KC) { return sizeof(float) * ((size_t)U * (4 * H + PAD) + (size_t)2 * CHUNK * (KC + PAD)); }

int check_common(int B, int T, int H, int ndir) {
    LAS_CHECK_ARG(B >= 1 && T >= 1, "lstm_rec: B=%d T=%d must be >= 1", B, T);
    LAS_CHECK_ARG(ndir == 1 || ndir == 2, "lstm_rec: ndir=%d must be 1 or 2", ndir);
    LAS_CHECK_ARG(H >= 32 && H % 32 == 0, "lstm_rec: hidden size %d must be a positive multiple of 32", H);
    const LasDeviceInfo* di = las_device_info();
    LAS_CHECK_ARG((H / U) * ndir <= di->num_sms, "lstm_rec: H=%d needs %d co-resident CTAs > %d SMs", H, (H / U) * ndir,
                  di->num_sms);
    return LAS_OK;
}

}  // namespace

extern "C" size_t las_lstm_rec_workspace_bytes(int B, int H, int ndir) {
    // barrier counters (256 B, kept apart) + dh/dc carried state
    return 256 + sizeof(float) * (size_t)2 * ndir * B * H;
}

extern "C" int las_lstm_rec_fwd_f32(float* gates, const float* w_hh, const int* lens, const float* drop_mask, float* out,
                                    float* hs_pad, float* cs_pad, int B, int T, int H, int ndir, void* ws, size_t ws_bytes,
                                    void* stream) {
    int rc = check_common(B, T, H, ndir);
    if (rc) return rc;
    LAS_CHECK_ARG(gates && w_hh && lens && hs_pad && cs_pad && ws, "lstm_rec_fwd: null pointer");
    if (ws_bytes < las_lstm_rec_workspace_bytes(B, H, ndir)) {
        las_set_error("lstm_rec_fwd: workspace %zu < %zu", ws_bytes, las_lstm_rec_workspace_bytes(B, H, ndir));
        return LAS_ERR_WORKSPACE;
    }
    rc = las_set_device_of(gates);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    RecArgs a{};
    a.gates = gates; a.w_hh = w_hh; a.lens = lens; a.mask = drop_mask; a.out = out; a.dout = nullptr;
    a.hs_pad = hs_pad; a.cs_pad = cs_pad; a.dstate = nullptr; a.ctr = (unsigned*)ws;
    a.B = B; a.T = T; a.H = H; a.ndir = ndir; a.NC = H / U; a.KC = pick_kc(H);
    size_t smem = fwd_smem(H, a.KC);
    LAS_CHECK_ARG(smem <= (size_t)las_device_info()->max_smem_optin, "lstm_rec_fwd: H=%d needs %zu B smem", H, smem);
    LAS_CUDA(cudaFuncSetAttribute(lstm_rec_fwd_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LAS_CUDA(cudaMemsetAsync(ws, 0, 256, st));
    LasProfScope prof(LAS_PROF_REC_FWD, stream, (double)T);
    void* args[] = {&a};
    LAS_CUDA(cudaLaunchCooperativeKernel((void*)lstm_rec_fwd_f32_kernel, dim3(a.NC, ndir), dim3(NTHREADS), args, smem, st));
    las_count_launch(1);
    return LAS_OK;
}

extern "C" int las_lstm_rec_bwd_f32(const float* dout, float* gates, const float* cs_pad, const float* w_hh, const int* lens,
                                    const float* drop_mask, int B, int T, int H, int ndir, void* ws, size_t ws_bytes,
                                    void* stream) {
    int rc = check_common(B, T, H, ndir);
    if (rc) return rc;
    LAS_CHECK_ARG(dout && gates && cs_pad && w_hh && lens && ws, "lstm_rec_bwd: null pointer");
    if (ws_bytes < las_lstm_rec_workspace_bytes(B, H, ndir)) {
        las_set_error("lstm_rec_bwd: workspace %zu < %zu", ws_bytes, las_lstm_rec_workspace_bytes(B, H, ndir));
        return LAS_ERR_WORKSPACE;
    }
    rc = las_set_device_of(gates);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    RecArgs a{};
    a.gates = gates; a.w_hh = w_hh; a.lens = lens; a.mask = drop_mask; a.out = nullptr; a.dout = dout;
    a.hs_pad = nullptr; a.cs_pad = const_cast<float*>(cs_pad); a.ctr = (unsigned*)ws;
    a.dstate = (float*)((char*)ws + 256);
    a.B = B; a.T = T; a.H = H; a.ndir = ndir; a.NC = H / U; a.KC = pick_kc(4 * H);
    size_t smem = bwd_smem(H, a.KC);
    LAS_CHECK_ARG(smem <= (size_t)las_device_info()->max_smem_optin, "lstm_rec_bwd: H=%d needs %zu B smem", H, smem);
    LAS_CUDA(cudaFuncSetAttribute(lstm_rec_bwd_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LAS_CUDA(cudaMemsetAsync(ws, 0, 256, st));
    LasProfScope prof(LAS_PROF_REC_BWD, stream, (double)T);
    void* args[] = {&a};
    LAS_CUDA(cudaLaunchCooperativeKernel((void*)lstm_rec_bwd_f32_kernel, dim3(a.NC, ndir), dim3(NTHREADS), args, smem, st));
    las_count_launch(1);
    return LAS_OK;
}
