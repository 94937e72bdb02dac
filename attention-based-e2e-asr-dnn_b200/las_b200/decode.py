"""Device-side transcripts of a greedy decode.

The reference builds every transcript on the host from the full logits tensor,
    batch_preds = [idx_to_str(pl.argmax(-1), VOCAB, SOS_IDX, EOS_IDX) for pl in pred_logits]        (src/infer.py:19-32,66)
one blocking device read per utterance.  Here the argmax the decoder fed back (model.spell.last_chars, device int32) -- or the argmax
of a logits tensor -- is cut on the device (every <sos> dropped, stop at the first <eos>) and ONE pinned copy of B * steps bytes plus
B lengths crosses the bus.

    logits, att = model(x, lx)                                  # eval mode
    strs = greedy_transcripts(model.spell.last_chars, VOCAB, SOS_IDX, EOS_IDX)      # == the reference's batch_preds
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch

from . import _lib
from ._lib import check, stream_ptr


def transcript_cut(chars: torch.Tensor, sos_idx: int, eos_idx: int, time_major: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
    """chars: (steps, B) [time_major, the layout of Speller.last_chars] or (B, steps) integer CUDA tensor of token ids < 256.
    Returns (tokens (B, steps) uint8, lengths (B,) int32) on the device: row b holds its kept tokens in [0, lengths[b])."""
    if not chars.is_cuda:
        raise RuntimeError('las_b200.decode.transcript_cut needs a CUDA tensor: there is no CPU fallback')
    c = chars.to(torch.int32)
    steps, B = (c.shape[0], c.shape[1]) if time_major else (c.shape[1], c.shape[0])
    ld_step, ld_b = (c.stride(0), c.stride(1)) if time_major else (c.stride(1), c.stride(0))
    out = torch.empty(B, steps, dtype=torch.uint8, device=c.device)
    lens = torch.empty(B, dtype=torch.int32, device=c.device)
    check(_lib.load().las_transcript_cut_i32(c.data_ptr(), int(ld_step), int(ld_b), int(B), int(steps), int(sos_idx), int(eos_idx),
                                             out.data_ptr(), lens.data_ptr(), stream_ptr()), 'transcript_cut')
    return out, lens


def greedy_transcripts(chars_or_logits: torch.Tensor, vocab: Sequence[str], sos_idx: int, eos_idx: int) -> List[str]:
    """Transcripts of a decoded batch, identical to [idx_to_str(pl.argmax(-1), vocab, sos, eos) for pl in pred_logits] (src/infer.py:66).
    Accepts Speller.last_chars ((steps, B) integer tensor) or pred_logits ((B, steps, V) float tensor)."""
    if chars_or_logits.dim() == 3:
        toks, lens = transcript_cut(chars_or_logits.argmax(-1), sos_idx, eos_idx, time_major=False)
    else:
        toks, lens = transcript_cut(chars_or_logits, sos_idx, eos_idx, time_major=True)
    B, steps = toks.shape
    host = torch.empty(4 * B + B * steps, dtype=torch.uint8, pin_memory=True)          # lengths and tokens in ONE copy
    packed = torch.cat([lens.view(torch.uint8), toks.reshape(-1)])
    host.copy_(packed, non_blocking=True)
    torch.cuda.current_stream(toks.device).synchronize()
    lens_h = host[:4 * B].view(torch.int32).tolist()
    rows = host[4 * B:].view(B, steps).numpy()
    return [''.join(vocab[int(t)] for t in rows[b, :lens_h[b]]) for b in range(B)]
