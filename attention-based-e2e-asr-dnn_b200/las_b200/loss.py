"""Fused, sync-free masked cross-entropy of the reference trainer (src/train.py:117-136).

    loss, ppl = masked_ce(pred_logits, y, ly, accu_grad=1)

`y` is the target tensor AFTER the trainer's `<sos>` strip (y[:, 1:]) on the device, `ly` the matching CPU length tensor
(ly - 1), exactly what src/train.py:117 has in hand.  When `pred_logits` has more steps than `y` (the dev loop decodes
CHR_MAX_STEPS steps and truncates, src/train.py:226-232) only the first y.size(1) steps enter the loss.  One kernel pass computes log-softmax, NLL, the length mask, the
masked mean and the gradient; `loss` and `ppl = exp(loss)` stay on the device (the reference's two `.item()` calls per batch,
src/train.py:149-150, become optional reads).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import check, stream_ptr


class MaskedCEFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, y, ly_dev, inv_denom):
        if not logits.is_cuda:
            raise RuntimeError('las_b200.masked_ce needs CUDA tensors: there is no CPU fallback')
        lib = _lib.load()
        B, S, V = logits.shape
        L = y.shape[1]                      # S > L: a longer (eval) decode, only the first L steps enter the loss (:226-232)
        if S < L:
            raise RuntimeError(f'masked_ce: {S} decoded steps < {L} target positions')
        lg = logits.detach().to(torch.float32).contiguous()
        yi = y.to(device=logits.device, dtype=torch.int32)
        if yi.stride(1) != 1:
            yi = yi.contiguous()
        out = torch.empty(2, dtype=torch.float32, device=logits.device)
        if ctx.needs_input_grad[0] and S != L:
            raise RuntimeError('masked_ce: the truncated (eval) form has no gradient')
        dl = torch.empty_like(lg) if ctx.needs_input_grad[0] else None
        ns = lib.las_masked_ce_scratch_floats(B, L)
        scratch = torch.empty(ns, dtype=torch.float32, device=logits.device)
        check(lib.las_masked_ce_f32(lg.data_ptr(), S * V, yi.data_ptr(), yi.stride(0), ly_dev.data_ptr(), B, L, V, float(inv_denom), out.data_ptr(),
                                    dl.data_ptr() if dl is not None else None, scratch.data_ptr(), ns, stream_ptr()), 'masked_ce')
        ctx.save_for_backward(dl)
        ctx.in_dtype = logits.dtype
        loss, ppl = out[0], out[1]
        ctx.mark_non_differentiable(ppl)
        return loss, ppl

    @staticmethod
    def backward(ctx, dloss, _dppl):
        (dl,) = ctx.saved_tensors
        return (dl * dloss).to(ctx.in_dtype), None, None, None


def masked_ce(pred_logits: torch.Tensor, y: torch.Tensor, ly, accu_grad: int = 1):
    """Returns (loss, ppl) as 0-dim device tensors.  ly: CPU int tensor / list of target lengths (already minus <sos>)."""
    L = y.shape[1]
    ly_cpu = torch.as_tensor(ly, dtype=torch.int64).cpu()
    n_nonpad = int(torch.clamp(ly_cpu, min=0, max=L).sum())           # == y_mask.sum() (src/train.py:125)
    if n_nonpad <= 0:
        raise RuntimeError('masked_ce: no non-padded target position')
    ly_dev = ly_cpu.to(device=pred_logits.device, dtype=torch.int32, non_blocking=True)
    return MaskedCEFunction.apply(pred_logits, y, ly_dev, 1.0 / (n_nonpad * accu_grad))
