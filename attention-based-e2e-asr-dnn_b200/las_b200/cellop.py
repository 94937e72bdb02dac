"""Standalone LSTM-cell pointwise autograd op (las_lstm_cell_{fwd,bwd}_f32) for callers that step the decoder cell
themselves (AutoRegDecoderLSTMCell.forward, the Rewriter of reference src/lmtrain.py:221-237)."""
from __future__ import annotations

import torch

from . import _lib
from ._lib import check, ptr, stream_ptr


class LSTMCellPointwise(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gates, c_prev, mask):
        if not gates.is_cuda:
            raise RuntimeError('las_b200 runs on CUDA tensors only (no CPU fallback)')
        gates = gates.float().contiguous().clone()       # activated in place below
        c_prev = c_prev.float().contiguous()
        B, H4 = gates.shape
        H = H4 // 4
        h = torch.empty(B, H, dtype=torch.float32, device=gates.device)
        c = torch.empty(B, H, dtype=torch.float32, device=gates.device)
        mask = mask.float().contiguous() if mask is not None else None
        check(_lib.load().las_lstm_cell_fwd_f32(gates.data_ptr(), c_prev.data_ptr(), ptr(mask), h.data_ptr(), c.data_ptr(), B, H,
                                                stream_ptr()), 'lstm_cell_fwd')
        ctx.save_for_backward(gates, c_prev, c, mask)
        return h, c

    @staticmethod
    def backward(ctx, dh, dc):
        gates, c_prev, c, mask = ctx.saved_tensors
        B, H = c.shape
        dG = gates.clone()
        dh = dh.float().contiguous() if dh is not None else torch.zeros_like(c)
        dc_io = dc.float().contiguous().clone() if dc is not None else torch.zeros_like(c)
        check(_lib.load().las_lstm_cell_bwd_f32(dG.data_ptr(), dh.data_ptr(), ptr(mask), c.data_ptr(), c_prev.data_ptr(),
                                                dc_io.data_ptr(), B, H, stream_ptr()), 'lstm_cell_bwd')
        return dG, dc_io, None


def lstm_cell_pointwise(gates, c_prev, mask=None):
    return LSTMCellPointwise.apply(gates, c_prev, mask)
