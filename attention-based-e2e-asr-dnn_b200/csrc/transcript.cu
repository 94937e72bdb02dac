// Device-side transcript cut: the reference turns each greedy index sequence into a string on the host,
//   idx_to_str(pl.argmax(-1), VOCAB, SOS_IDX, EOS_IDX)   (src/infer.py:19-32, called per utterance at :66; src/train.py:405-419 for
//   the dev-set edit distance) : skip every <sos>, stop at the first <eos>, map the rest through the vocabulary,
// after copying the whole (B, steps, V) logits tensor to the host one utterance at a time.  Here the argmax the decoder already fed
// back (chars, (steps, B) int32, device) is compacted on the device -- <sos> dropped, cut at the first <eos> -- into (B, steps) bytes
// plus a length per utterance, so ONE small D2H copy carries every transcript of the batch (B * steps bytes instead of B * steps * V
// floats), and the host only maps bytes to characters.  One warp per utterance, ballot / popc prefix compaction.
#include "las_common.cuh"
#include "las_b200.h"

namespace {
__global__ void __launch_bounds__(256) transcript_cut_kernel(const int* __restrict__ chars, long long ld_step, long long ld_b, int B, int steps,
                                                             int sos, int eos, unsigned char* __restrict__ out, int* __restrict__ lens) {
    const int b = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= B) return;
    int n = 0;
    bool done = false;
    for (int t0 = 0; t0 < steps && !done; t0 += 32) {
        const int t = t0 + lane;
        const int c = t < steps ? chars[(long long)t * ld_step + (long long)b * ld_b] : eos;
        const unsigned m_eos = __ballot_sync(0xffffffffu, t < steps && c == eos);
        const unsigned before = m_eos ? ((1u << (__ffs(m_eos) - 1)) - 1u) : 0xffffffffu;      // lanes ahead of the first <eos>
        const bool keep = t < steps && c != sos && ((before >> lane) & 1u);
        const unsigned m_keep = __ballot_sync(0xffffffffu, keep);
        if (keep) out[(long long)b * steps + n + __popc(m_keep & ((1u << lane) - 1u))] = (unsigned char)c;
        n += __popc(m_keep);
        done = m_eos != 0u;
    }
    if (lane == 0) lens[b] = n;
}
}  // namespace

extern "C" int las_transcript_cut_i32(const int* chars, long long ld_step, long long ld_b, int B, int steps, int sos_idx, int eos_idx,
                                      unsigned char* out, int* lens, void* stream) {
    LAS_CHECK_ARG(chars && out && lens && B >= 1 && steps >= 1, "transcript_cut: bad arguments");
    int rc = las_set_device_of(out);
    if (rc) return rc;
    transcript_cut_kernel<<<ceil_div(B, 8), 256, 0, (cudaStream_t)stream>>>(chars, ld_step, ld_b, B, steps, sos_idx, eos_idx, out, lens);
    LAS_LAUNCH_CHECK();
    return LAS_OK;
}
