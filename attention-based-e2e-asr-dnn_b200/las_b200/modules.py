"""Block-level modules with the reference's API (src/modules.py): LockedLSTM, pyramLockedLSTM,
AutoRegDecoderLSTMCell.  Same class names, constructor kwargs, forward signatures, parameter names/shapes
(state_dict keys are a frozen contract) and public attributes.

nn.LSTM / nn.LSTMCell / nn.Dropout objects are kept ONLY as parameter containers so that state_dict keys match the
reference; their forward() is never called -- the arithmetic runs in liblas_b200.so (hand-written sm_100a kernels).
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn as nn

from . import functional as LF

# ---- test hook: inject dropout masks instead of drawing them (parity with recorded reference masks) ----
_MASK_OVERRIDE = {'locked': None, 'drop': None, 'coins': None}


def set_mask_override(locked: Optional[List[torch.Tensor]] = None, drop: Optional[List[torch.Tensor]] = None,
                      coins: Optional[List[float]] = None):
    """locked: list of (B,1,F) masks consumed in layer order; drop: list of (B,hid) masks consumed in the reference's
    nn.Dropout call order (cell0, cell1 per step); coins: the raw torch.rand(1) teacher-forcing draws, one per step
    t >= 1.  Pass None to clear."""
    _MASK_OVERRIDE['locked'] = list(locked) if locked is not None else None
    _MASK_OVERRIDE['drop'] = list(drop) if drop is not None else None
    _MASK_OVERRIDE['coins'] = list(coins) if coins is not None else None


def _locked_mask(x_like_B, F_, p, training, device):
    """Reference: x.new_empty(B, 1, F).bernoulli_(1 - p).div_(1 - p) (src/modules.py:61-63, :149-152)."""
    if (not training) or (not p):
        return None
    if _MASK_OVERRIDE['locked'] is not None:
        return _MASK_OVERRIDE['locked'].pop(0).to(device=device, dtype=torch.float32)
    return torch.empty(x_like_B, 1, F_, dtype=torch.float32, device=device).bernoulli_(1 - p).div_(1 - p)


def _check_lengths(lx: torch.Tensor) -> torch.Tensor:
    """pack_padded_sequence's contract (src/modules.py:78): a CPU int64 1-D tensor of positive lengths."""
    if not isinstance(lx, torch.Tensor):
        lx = torch.as_tensor(lx, dtype=torch.int64)
    if lx.is_cuda:
        raise RuntimeError("'lengths' argument should be a 1D CPU int64 tensor, but got 1D cuda:0 Long tensor")
    lx = lx.to(torch.int64)
    if lx.numel() == 0 or int(lx.min()) <= 0:
        raise RuntimeError("Length of all samples has to be greater than 0, but found an element in 'lengths' that is <= 0")
    return lx


def _lstm_weights(lstm: nn.LSTM):
    ws = [lstm.weight_ih_l0, lstm.weight_hh_l0, lstm.bias_ih_l0, lstm.bias_hh_l0]
    if lstm.bidirectional:
        ws += [lstm.weight_ih_l0_reverse, lstm.weight_hh_l0_reverse, lstm.bias_ih_l0_reverse, lstm.bias_hh_l0_reverse]
    return ws


def _first_layer_hint(block, lx):
    """next_layer hint for the layer that feeds `block` (a pyramLockedLSTM: its first layer halves the time axis)."""
    if block is None or not isinstance(block, pyramLockedLSTM) or len(block.plstms) == 0:
        return None
    Tn = int((lx // 2).max())
    return dict(weights=_lstm_weights(block.plstms[0]), pyramid=True, T=Tn) if Tn > 0 else None


class LockedLSTM(nn.Module):
    """Stack of 1-layer (Bi)LSTMs with locked dropout -- reference src/modules.py:11-85."""

    def __init__(self, lstm_input_dim: int = 15, uniform_hid_dim: int = 256, lstm_layers: int = 1, bidirectional: bool = True,
                 init_dropout: float = 0.2, mid_dropout: float = 0.3):
        super().__init__()
        self.lstm_input_dim = lstm_input_dim
        self.uniform_hid_dim = uniform_hid_dim
        self.lstm_layers = lstm_layers
        self.bidirectional = bidirectional
        self.init_dropout = init_dropout
        self.mid_dropout = mid_dropout
        self.lstms = nn.ModuleList([
            nn.LSTM(input_size=(self.lstm_input_dim if i == 0 else self.uniform_hid_dim * (int(self.bidirectional) + 1)),
                    hidden_size=self.uniform_hid_dim, num_layers=1, batch_first=True, dropout=0,
                    bidirectional=self.bidirectional)
            for i in range(self.lstm_layers)])

    def forward(self, x, lx):
        lx = _check_lengths(lx)
        for i, lstm in enumerate(self.lstms):
            p = self.mid_dropout if i else self.init_dropout
            T = int(lx.max())
            if T > x.size(1):
                raise RuntimeError(f'Expected sequence length to be larger than 0 and at most {x.size(1)}, got {T}')
            lens_dev = lx.to(device=x.device, dtype=torch.int32, non_blocking=True)
            F_ = self.uniform_hid_dim * (int(self.bidirectional) + 1)
            mask = _locked_mask(x.size(0), F_, p, self.training, x.device)
            # what consumes this layer's output (lets it project the consumer's gates beside its own recurrence, LF "Forward pipelining")
            if i + 1 < len(self.lstms):
                nxt = dict(weights=_lstm_weights(self.lstms[i + 1]), pyramid=False, T=T)
            else:
                nxt = _first_layer_hint(self.__dict__.get('_las_next'), lx)
            x = LF.lstm_layer(x, lens_dev, T, False, mask, _lstm_weights(lstm), nxt)
        return x, lx.clone()


class pyramLockedLSTM(nn.Module):
    """Pyramidal BiLSTM stack (frame-pair concat, 2x time reduction per layer) with locked dropout -- reference
    src/modules.py:89-194.  The odd-frame drop (:171-181), lx // 2 (:183) and reshape (:185) are addressing only."""

    def __init__(self, plstm_input_dim: int = 512, uniform_hid_dim: int = 256, plstm_layers: int = 3, bidirectional: bool = True,
                 mid_dropout: float = 0.2, final_dropout: float = 0.2):
        super().__init__()
        self.plstm_input_dim = plstm_input_dim
        self.uniform_hid_dim = uniform_hid_dim
        self.plstm_layers = plstm_layers
        self.bidirectional = bidirectional
        self.mid_dropout = mid_dropout
        self.final_dropout = final_dropout
        self.dims = [2 * self.uniform_hid_dim * (int(self.bidirectional) + 1) for _ in range(self.plstm_layers)]
        self.dims[0] = 2 * self.plstm_input_dim
        self.plstms = nn.ModuleList([
            nn.LSTM(input_size=self.dims[i], hidden_size=self.uniform_hid_dim, num_layers=1, batch_first=True, dropout=0,
                    bidirectional=self.bidirectional)
            for i in range(self.plstm_layers)])

    def forward(self, x, lx):
        lx = _check_lengths(lx)
        for i, plstm in enumerate(self.plstms):
            p = self.mid_dropout if i < self.plstm_layers - 1 else self.final_dropout
            lx = lx // 2
            _check_lengths(lx)
            T = int(lx.max())
            if 2 * T > x.size(1):
                raise RuntimeError(f'Expected sequence length to be larger than 0 and at most {x.size(1) // 2}, got {T}')
            lens_dev = lx.to(device=x.device, dtype=torch.int32, non_blocking=True)
            F_ = self.uniform_hid_dim * (int(self.bidirectional) + 1)
            mask = _locked_mask(x.size(0), F_, p, self.training, x.device)
            nxt = None
            if i + 1 < len(self.plstms) and int((lx // 2).max()) > 0:
                nxt = dict(weights=_lstm_weights(self.plstms[i + 1]), pyramid=True, T=int((lx // 2).max()))
            x = LF.lstm_layer(x, lens_dev, T, True, mask, _lstm_weights(plstm), nxt)
        return x, lx.clone()


class AutoRegDecoderLSTMCell(nn.Module):
    """Two stacked LSTM cells of the Speller -- reference src/modules.py:302-365.  Parameter container for the fused
    decoder loop (las_speller_*); `forward` gives the reference's single-step API for external callers."""

    def __init__(self, att_proj_dim: int = 128, dec_emb_dim: int = 256, dec_hid_dim: int = 512, dec_out_dim: int = 128,
                 dec_mid_dropout: float = 0.2):
        super().__init__()
        self.att_proj_dim = att_proj_dim
        self.dec_emb_dim = dec_emb_dim
        self.dec_hid_dim = dec_hid_dim
        self.dec_out_dim = dec_out_dim
        self.dec_mid_dropout = dec_mid_dropout
        self.lstms = nn.ModuleList([
            nn.LSTMCell(input_size=self.att_proj_dim + self.dec_emb_dim, hidden_size=self.dec_hid_dim),
            nn.LSTMCell(input_size=self.dec_hid_dim, hidden_size=self.dec_out_dim)])
        self.dropout = nn.Dropout(self.dec_mid_dropout)

    def draw_dropout_mask(self, B, hid, device):
        """nn.Dropout(p)(h) == h * mask with mask = F.dropout(ones, p, True) (src/modules.py:356): same RNG consumption."""
        p = self.dec_mid_dropout
        if (not self.training) or (not p):
            return None
        if _MASK_OVERRIDE['drop'] is not None:
            return _MASK_OVERRIDE['drop'].pop(0).to(device=device, dtype=torch.float32)
        return torch.nn.functional.dropout(torch.ones(B, hid, dtype=torch.float32, device=device), p, True)

    def forward(self, prev_e, prev_c, prev_h):
        from .functional import linear
        from .cellop import lstm_cell_pointwise
        prev_ec = torch.cat([prev_e, prev_c], dim=1)
        for i in range(len(self.lstms)):
            cell = self.lstms[i]
            h_prev, c_prev = prev_h[i]
            # one GEMM over the packed row [x | h] against [W_ih | W_hh]; both biases fused in the epilogue
            gates = linear(torch.cat([prev_ec, h_prev], dim=1), torch.cat([cell.weight_ih, cell.weight_hh], dim=1),
                           cell.bias_ih, cell.bias_hh)
            mask = self.draw_dropout_mask(gates.size(0), cell.hidden_size, gates.device)
            h, c = lstm_cell_pointwise(gates, c_prev, mask)
            prev_h[i] = (h, c)
            if i == 0:
                prev_ec = prev_h[i][0]
        return prev_h
