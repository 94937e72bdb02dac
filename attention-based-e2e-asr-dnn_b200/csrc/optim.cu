// Fused GradScaler.unscale_ + clip_grad_norm_ + AdamW(amsgrad) over a table of tensors.
//
// Replaces, in two launches, the ~15 foreach launches of reference src/train.py:165-183:
//   scaler.unscale_(opt)              g *= 1/scale ; found_inf = any(!isfinite(g))     (torch amp/grad_scaler.py)
//   clip_grad_norm_(params, 5.0)      total = ||(||g_i||)||_2 ; coef = min(max_norm/(total+1e-6), 1) ; g *= coef
//   scaler.step(AdamW amsgrad)        skipped entirely when found_inf, else (torch optim/adam.py single-tensor):
//       p *= 1 - lr*wd ; m = lerp(m, g, 1-b1) ; v = b2*v + (1-b2) g^2 ; vmax = max(vmax, v)
//       p -= (lr / bias_c1) * m / (sqrt(vmax)/sqrt(bias_c2) + eps)
// HBM bound: reads g,p,m,v,vmax and writes p,m,v,vmax once (36 B/element); pass 1 reads g once more.
#include "las_common.cuh"
#include "las_b200.h"

namespace {
constexpr int NT = 256;

__global__ void __launch_bounds__(NT) adam_pass1_kernel(const LasAdamTensor* __restrict__ table, const LasAdamChunk* __restrict__ chunks,
                                                        float inv_scale, float* __restrict__ partial, float* __restrict__ status) {
    const LasAdamChunk ck = chunks[blockIdx.x];
    const LasAdamTensor tn = table[ck.tensor];
    const long long n = min((long long)LAS_ADAM_CHUNK, tn.numel - ck.offset);
    const float* g = tn.g + ck.offset;
    float acc = 0.f;
    bool bad = false;
    for (long long i = threadIdx.x; i < n; i += NT) {
        const float v = g[i] * inv_scale;
        if (!isfinite(v)) bad = true;
        acc = fmaf(v, v, acc);
    }
    __shared__ float red[NT / 32];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    const int anybad = __syncthreads_or(bad ? 1 : 0);
    if (threadIdx.x == 0) {
        float r = 0.f;
#pragma unroll
        for (int i = 0; i < NT / 32; ++i) r += red[i];
        partial[blockIdx.x] = r;
        if (anybad) status[0] = 1.f;   // benign race: every writer stores the same value
    }
}

__global__ void __launch_bounds__(NT) adam_pass2_kernel(const LasAdamTensor* __restrict__ table, const LasAdamChunk* __restrict__ chunks,
                                                        int n_chunks, float one_minus_b1, float beta2, float one_minus_b2, float eps,
                                                        float decay, float inv_scale, float max_norm, int amsgrad,
                                                        const float* __restrict__ partial, float* __restrict__ status) {
    // every block re-reduces the per-chunk partial sums in the same order: deterministic, no extra launch
    __shared__ float red[NT / 32];
    __shared__ float s_coef;
    float acc = 0.f;
    for (int i = threadIdx.x; i < n_chunks; i += NT) acc += partial[i];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float r = 0.f;
#pragma unroll
        for (int i = 0; i < NT / 32; ++i) r += red[i];
        const float total = sqrtf(r);
        s_coef = fminf(max_norm / (total + 1e-6f), 1.0f);
        if (blockIdx.x == 0) status[1] = total;
    }
    __syncthreads();
    if (status[0] != 0.f) return;          // found_inf: GradScaler.step skips the optimizer step
    const float gmul = inv_scale * ((max_norm > 0.f) ? s_coef : 1.0f);
    const LasAdamChunk ck = chunks[blockIdx.x];
    const LasAdamTensor tn = table[ck.tensor];
    const long long n = min((long long)LAS_ADAM_CHUNK, tn.numel - ck.offset);
    const float* g = tn.g + ck.offset;
    float* p = tn.p + ck.offset;
    float* m = tn.m + ck.offset;
    float* v = tn.v + ck.offset;
    float* vm = tn.vmax ? tn.vmax + ck.offset : nullptr;
    const float step_size = tn.step_size;
    for (long long i = threadIdx.x; i < n; i += NT) {
        const float gi = g[i] * gmul;
        float pi = p[i] * decay;
        float mi = m[i];
        mi = mi + (gi - mi) * one_minus_b1;
        float vi = v[i] * beta2 + one_minus_b2 * (gi * gi);
        float dv = vi;
        if (amsgrad) {
            dv = fmaxf(vm[i], vi);
            vm[i] = dv;
        }
        const float denom = sqrtf(dv) / tn.bias_c2_sqrt + eps;
        pi -= step_size * (mi / denom);
        p[i] = pi; m[i] = mi; v[i] = vi;
    }
}
}  // namespace

extern "C" int las_adamw_amsgrad_fused(const LasAdamTensor* table, int n_tensors, const LasAdamChunk* chunks, int n_chunks, double lr,
                                       double beta1_d, double beta2_d, float eps, double weight_decay, float inv_scale, float max_norm,
                                       int amsgrad, float* scratch, float* status, void* stream) {
    LAS_CHECK_ARG(table && chunks && scratch && status && n_tensors >= 1 && n_chunks >= 1, "adamw: bad arguments");
    int rc = las_set_device_of(table);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    LasProfScope prof(LAS_PROF_ADAM, stream, 0.0);
    LAS_CUDA(cudaMemsetAsync(status, 0, 2 * sizeof(float), st));
    adam_pass1_kernel<<<n_chunks, NT, 0, st>>>(table, chunks, inv_scale, scratch, status);
    LAS_LAUNCH_CHECK();
    // scalar constants are formed in double on the host and rounded once, like torch's Python-float arithmetic
    const float one_minus_b1 = (float)(1.0 - (double)beta1_d), one_minus_b2 = (float)(1.0 - (double)beta2_d);
    const float decay = (float)(1.0 - (double)lr * (double)weight_decay);
    adam_pass2_kernel<<<n_chunks, NT, 0, st>>>(table, chunks, n_chunks, one_minus_b1, (float)beta2_d, one_minus_b2, eps, decay, inv_scale,
                                               max_norm, amsgrad, scratch, status);
    LAS_LAUNCH_CHECK();
    return LAS_OK;
}
