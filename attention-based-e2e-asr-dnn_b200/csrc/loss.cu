// Fused masked cross-entropy of the trainer (reference src/train.py:117-136):
//   y_mask[b,t] = t < ly[b] ;  loss = sum(CE_none(logits, y) * y_mask) / (n_nonpad * accu_grad) ;  ppl = exp(loss)
// One pass over the (B*L, V) logits: log-softmax, NLL, mask, the gradient (softmax - onehot) * mask * inv_denom, and a
// deterministic two-stage sum -- no host synchronisation (n_nonpad comes from the CPU length tensor the loader provides).
#include "las_common.cuh"
#include "las_b200.h"
#include <float.h>

namespace {

constexpr int ROWS_PER_BLOCK = 8;     // one warp per (b, t) row

__global__ void __launch_bounds__(256) masked_ce_kernel(const float* __restrict__ logits, long long ld_b, const int* __restrict__ y, long long ld_y,
                                                        const int* __restrict__ ly, int B, int L, int V, float inv_denom,
                                                        float* __restrict__ dlogits, float* __restrict__ partial) {
    __shared__ float red[ROWS_PER_BLOCK];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * ROWS_PER_BLOCK + w;
    float nll = 0.f;
    if (row < (long long)B * L) {
        const int b = (int)(row / L), t = (int)(row - (long long)b * L);
        const bool on = t < ly[b];
        const float* lr = logits + (long long)b * ld_b + (long long)t * V;      // ld_b > L*V: logits of a longer decode, truncated to L steps
        const int tgt = y[(long long)b * ld_y + t];
        float mx = -FLT_MAX;
        for (int v = lane; v < V; v += 32) mx = fmaxf(mx, lr[v]);
        mx = warp_max(mx);
        float sum = 0.f;
        for (int v = lane; v < V; v += 32) sum += expf(lr[v] - mx);
        sum = warp_sum(sum);
        const float lse = logf(sum) + mx;
        // a non-masked target outside [0, V): nn.CrossEntropyLoss raises (host-side check, src/train.py:131-136); a sync-free kernel
        // cannot, so the loss, the perplexity and this row's gradient become NaN -- loud (GradScaler skips the step, the trainer's
        // printed loss is nan) instead of a silently wrong gradient
        const bool bad = on && (tgt < 0 || tgt >= V);
        if (on) nll = bad ? __int_as_float(0x7fc00000) : lse - lr[tgt];
        if (dlogits) {
            float* dr = dlogits + row * V;
            for (int v = lane; v < V; v += 32)
                dr[v] = bad ? __int_as_float(0x7fc00000) : (on ? (expf(lr[v] - lse) - (v == tgt ? 1.f : 0.f)) * inv_denom : 0.f);
        }
    }
    if (lane == 0) red[w] = nll;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < ROWS_PER_BLOCK; ++i) s += red[i];
        partial[blockIdx.x] = s;
    }
}

__global__ void __launch_bounds__(256) masked_ce_final_kernel(const float* __restrict__ partial, int n, float inv_denom, float* __restrict__ out) {
    __shared__ float red[256];
    float s = 0.f;
    for (int i = threadIdx.x; i < n; i += 256) s += partial[i];     // fixed order per thread, fixed tree below: deterministic
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const float loss = red[0] * inv_denom;
        out[0] = loss;
        out[1] = expf(loss);        // perplexity (src/train.py:139)
    }
}

}  // namespace

extern "C" size_t las_masked_ce_scratch_floats(int B, int L) { return (size_t)ceil_div64((long long)B * L, ROWS_PER_BLOCK) + 4; }

extern "C" int las_masked_ce_f32(const float* logits, long long ld_b, const int* y, long long ld_y, const int* ly_dev, int B, int L, int V, float inv_denom,
                                 float* loss_ppl_out, float* dlogits, float* scratch, size_t scratch_floats, void* stream) {
    LAS_CHECK_ARG(logits && y && ly_dev && loss_ppl_out && scratch && B >= 1 && L >= 1 && V >= 2 && ld_b >= (long long)L * V, "masked_ce: bad arguments");
    LAS_CHECK_ARG(scratch_floats >= las_masked_ce_scratch_floats(B, L), "masked_ce: scratch too small");
    int rc = las_set_device_of(logits);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int nblk = (int)ceil_div64((long long)B * L, ROWS_PER_BLOCK);
    masked_ce_kernel<<<nblk, 256, 0, st>>>(logits, ld_b, y, ld_y, ly_dev, B, L, V, inv_denom, dlogits, scratch);
    LAS_LAUNCH_CHECK();
    masked_ce_final_kernel<<<1, 256, 0, st>>>(scratch, nblk, inv_denom, loss_ppl_out);
    LAS_LAUNCH_CHECK();
    return LAS_OK;
}
