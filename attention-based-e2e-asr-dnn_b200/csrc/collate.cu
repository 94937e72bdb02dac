// Device-side collate + SpecAugment of the training loader (reference src/utils.py:95-128, :82-84).
// The reference pads the length-sorted MFCC list on the host (pad_sequence), optionally applies torchaudio's
// FrequencyMasking(6) / TimeMasking(200) to the padded (B, F, T) batch -- ONE mask interval per axis for the whole batch
// (torchaudio functional.mask_along_axis) -- and the trainer then copies the padded batch to the GPU.  Here the ragged
// frames go up once (no padding bytes over the bus) and one kernel writes the padded, masked (B, T, F) batch.
#include "las_common.cuh"
#include "las_b200.h"

namespace {

__global__ void __launch_bounds__(256) collate_kernel(const float* __restrict__ frames, const long long* __restrict__ offsets,
                                                      const int* __restrict__ lens, int B, int T, int F, float pad_value, int f_lo, int f_hi,
                                                      int t_lo, int t_hi, float mask_value, float* __restrict__ out) {
    const long long total = (long long)B * T * F;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int f = (int)(i % F);
        const long long bt = i / F;
        const int t = (int)(bt % T), b = (int)(bt / T);
        float v = (t < lens[b]) ? frames[(offsets[b] + t) * F + f] : pad_value;
        if ((f >= f_lo && f < f_hi) || (t >= t_lo && t < t_hi)) v = mask_value;      // masked_fill over the padded batch
        out[i] = v;
    }
}

}  // namespace

extern "C" int las_collate_specaug_f32(const float* frames, const long long* offsets, const int* lens, int B, int T, int F, float pad_value,
                                       int f_lo, int f_hi, int t_lo, int t_hi, float mask_value, float* out, void* stream) {
    LAS_CHECK_ARG(frames && offsets && lens && out && B >= 1 && T >= 1 && F >= 1, "collate: bad arguments");
    int rc = las_set_device_of(out);
    if (rc) return rc;
    const long long total = (long long)B * T * F;
    long long blocks = (total + 255) / 256;
    const int cap = las_device_info()->num_sms * 8;
    if (blocks > cap) blocks = cap;
    collate_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(frames, offsets, lens, B, T, F, pad_value, f_lo, f_hi, t_lo, t_hi, mask_value, out);
    LAS_LAUNCH_CHECK();
    return LAS_OK;
}
