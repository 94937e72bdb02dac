"""Model-level modules with the reference's API (src/models.py): Listener, MultiheadCrossAttention, Speller,
ListenAttendSpell.  Same constructor kwargs (splatted from config/*.yml), forward signatures, returns, parameter
names/shapes (state_dict contract) and the public attributes train.py / infer.py / lmtrain.py touch.
"""
from __future__ import annotations

import math
import os
from typing import Optional

import torch
import torch.nn as nn

from . import functional as LF
from .modules import AutoRegDecoderLSTMCell, LockedLSTM, pyramLockedLSTM, _MASK_OVERRIDE


class Listener(nn.Module):
    """reference src/models.py:16-66"""

    def __init__(self, input_dim: int = 15, uniform_hid_dim: int = 256, lstm_layers: int = 1, plstm_layers: int = 3,
                 bidirectional: bool = True, init_dropout: float = 0.2, mid_dropout: float = 0.3, final_dropout: float = 0.4):
        super().__init__()
        self.input_dim = input_dim
        self.uniform_hid_dim = uniform_hid_dim
        self.lstm_layers = lstm_layers
        self.plstm_layers = plstm_layers
        self.bidirectional = bidirectional
        # shadow attributes mutated by Trainer.dropout_step (src/train.py:467-474); as in the reference the blocks read
        # their own copies, so changing these has no effect
        self.init_dropout = init_dropout
        self.mid_dropout = mid_dropout
        self.final_dropout = final_dropout
        self.base = LockedLSTM(lstm_input_dim=self.input_dim, uniform_hid_dim=self.uniform_hid_dim, lstm_layers=self.lstm_layers,
                               bidirectional=self.bidirectional, init_dropout=self.init_dropout, mid_dropout=self.mid_dropout)
        self.pyramid = pyramLockedLSTM(plstm_input_dim=(int(self.bidirectional) + 1) * self.uniform_hid_dim,
                                       uniform_hid_dim=self.uniform_hid_dim, plstm_layers=self.plstm_layers,
                                       bidirectional=self.bidirectional, mid_dropout=self.mid_dropout,
                                       final_dropout=self.final_dropout)

        # the base stack's last layer feeds the pyramid's first: told through the instance dict so that no submodule gets registered twice
        self.base.__dict__['_las_next'] = self.pyramid

    def forward(self, x, lx):
        return self.pyramid(*self.base(x, lx))


class MultiheadCrossAttention(nn.Module):
    """reference src/models.py:70-192.  `wrapup_encodings` caches keys/values/masks on self exactly like the reference
    (same shapes: keys (B,h,d,T) view, values (B,h,T,d) view, masks (B,h,1,T) bool); `forward` is the single fused
    attention step.  final_map exists (state_dict contract) and, as in the reference, is never used."""

    def __init__(self, enc_out_dim: int = 512, dec_out_dim: int = 128, proj_dim: int = 128, heads: int = 4, dropout: float = 0.1):
        super().__init__()
        assert proj_dim % heads == 0
        self.enc_out_dim = enc_out_dim
        self.dec_out_dim = dec_out_dim
        self.proj_dim = proj_dim
        self.heads = heads
        self.dims_per_head = self.proj_dim // self.heads
        self.norm_factor = 1 / math.sqrt(self.dims_per_head)
        self.key_map = nn.Linear(self.enc_out_dim, self.proj_dim)
        self.value_map = nn.Linear(self.enc_out_dim, self.proj_dim)
        self.query_map = nn.Linear(self.dec_out_dim, self.proj_dim)
        self.final_map = nn.Linear(self.proj_dim, self.proj_dim)
        self.softmax = nn.Softmax(dim=-1)
        self.dropout = dropout

    @staticmethod
    def build_pad_masks(enc_l):
        max_len = enc_l.max()
        return (torch.arange(0, max_len, dtype=torch.int64).unsqueeze(0) >= enc_l.unsqueeze(1))

    def project_memory(self, enc_h, enc_l):
        """K, V as (B, T, P) row-major fp32 (the kernels' layout) and device lengths.  No mask tensor, no H2D of a mask:
        the kernels mask from lengths."""
        B, T, _ = enc_h.shape
        self._K = LF.linear(enc_h, self.key_map.weight, self.key_map.bias)
        self._V = LF.linear(enc_h, self.value_map.weight, self.value_map.bias)
        self._lens_dev = torch.as_tensor(enc_l, dtype=torch.int64).to(device=enc_h.device, dtype=torch.int32, non_blocking=True)
        return self._K, self._V, self._lens_dev

    def wrapup_encodings(self, enc_h, enc_l):
        B, T, _ = enc_h.shape
        K, V, _ = self.project_memory(enc_h, enc_l)
        self.keys = K.view(B, T, self.heads, self.dims_per_head).transpose(1, 2).transpose(-2, -1)
        self.values = V.view(B, T, self.heads, self.dims_per_head).transpose(1, 2)
        mask = self.build_pad_masks(torch.as_tensor(enc_l, dtype=torch.int64).cpu())
        self.masks = mask[:, None, None, :].expand((B, self.heads, 1, T)).to(enc_h.device)

    def forward(self, dec_h, return_wgts: bool = False, init_wgts_mask: torch.Tensor = None):
        B = dec_h.size(0)
        q = LF.linear(dec_h, self.query_map.weight, self.query_map.bias)
        self.queries = q.view(B, self.heads, self.dims_per_head).unsqueeze(2)
        fmask = None
        if init_wgts_mask is not None:          # (B, heads, 1, T) prior (reference :177-181), usually an expanded view
            T = self._K.shape[1]
            fmask = init_wgts_mask.to(device=q.device, dtype=torch.float32).expand(B, self.heads, 1, T).reshape(B * self.heads, T)
        ctx, w = LF.attn_step(q, self._K, self._V, self._lens_dev, self.heads, fmask)
        wgts = w.view(B, self.heads, 1, -1)
        if init_wgts_mask is not None:
            return ctx, wgts.detach()           # the reference returns the detached pre-prior weights (:178,188)
        return (ctx, wgts) if return_wgts else ctx


class Speller(nn.Module):
    """reference src/models.py:197-386.  forward() runs the whole decoder loop in one C call per direction
    (las_speller_fwd_f32 / las_speller_bwd_f32): no per-step Python, no per-step host sync."""

    def __init__(self, enc_out_dim: int = 512, att_proj_dim: int = 128, att_heads: int = 4, att_dropout: float = 0.2,
                 dec_vocab_size: int = 30, dec_emb_dim: int = 256, dec_emb_dropout: float = 0.5, dec_lstm_hid_dim: int = 512,
                 dec_lstm_out_dim: int = 128, dec_lstm_dropout: float = 0.2, CHR_MAX_STEPS: int = 600, CHR_PAD_IDX: int = 29,
                 CHR_SOS_IDX: int = 0, USE_GREEDY: bool = True):
        super().__init__()
        self.enc_out_dim = enc_out_dim
        self.att_proj_dim = att_proj_dim
        self.att_heads = att_heads
        self.att_dropout = att_dropout
        self.dec_vocab_size = dec_vocab_size
        self.dec_emb_dim = dec_emb_dim
        self.dec_emb_dropout = dec_emb_dropout
        self.dec_lstm_hid_dim = dec_lstm_hid_dim
        self.dec_lstm_out_dim = dec_lstm_out_dim
        self.dec_lstm_dropout = dec_lstm_dropout
        self.CHR_MAX_STEPS = CHR_MAX_STEPS
        self.CHR_PAD_IDX = CHR_PAD_IDX
        self.CHR_SOS_IDX = CHR_SOS_IDX
        self.USE_GREEDY = USE_GREEDY

        self.attention = MultiheadCrossAttention(enc_out_dim=self.enc_out_dim, dec_out_dim=self.dec_lstm_out_dim,
                                                 proj_dim=self.att_proj_dim, heads=self.att_heads, dropout=self.att_dropout)
        self.char_emb = nn.Embedding(num_embeddings=self.dec_vocab_size, embedding_dim=self.dec_emb_dim,
                                     padding_idx=self.CHR_PAD_IDX)
        self.lstms = AutoRegDecoderLSTMCell(att_proj_dim=self.att_proj_dim, dec_emb_dim=self.dec_emb_dim,
                                            dec_hid_dim=self.dec_lstm_hid_dim, dec_out_dim=self.dec_lstm_out_dim,
                                            dec_mid_dropout=self.dec_lstm_dropout)
        self.init_query = nn.Parameter(torch.rand((1, self.dec_lstm_out_dim)), requires_grad=True)
        # unregistered, never trained, always zero (reference :275-281) -- kept for attribute parity
        self.init_hiddens = [(nn.Parameter(torch.zeros((1, self.dec_lstm_hid_dim)), requires_grad=True),
                              nn.Parameter(torch.zeros((1, self.dec_lstm_hid_dim)), requires_grad=True)),
                             (nn.Parameter(torch.zeros((1, self.dec_lstm_out_dim)), requires_grad=True),
                              nn.Parameter(torch.zeros((1, self.dec_lstm_out_dim)), requires_grad=True))]
        self.cls = nn.Linear(self.dec_emb_dim, self.dec_vocab_size)
        self.cls.weight = self.char_emb.weight            # weight tying (:287)
        if self.dec_emb_dim != 2 * self.att_proj_dim:
            raise ValueError(f'dec_emb_dim ({self.dec_emb_dim}) must equal 2*att_proj_dim ({2 * self.att_proj_dim}): the tied '
                             'classifier consumes cat[q_proj, context] (reference src/models.py:285-287,371)')

    def _decoder_masks(self, steps, B, device):
        """nn.Dropout masks of the two cells for every step (src/modules.py:356).  Default: two batched draws.
        LAS_EXACT_RNG=1 reproduces the reference's RNG consumption order (one F.dropout call per cell per step)."""
        p = self.lstms.dec_mid_dropout
        if (not self.training) or (not p):
            return None, None
        DH, DO = self.dec_lstm_hid_dim, self.dec_lstm_out_dim
        if _MASK_OVERRIDE['drop'] is not None or os.environ.get('LAS_EXACT_RNG', '0') == '1':
            m0, m1 = [], []
            for _ in range(steps):
                m0.append(self.lstms.draw_dropout_mask(B, DH, device))
                m1.append(self.lstms.draw_dropout_mask(B, DO, device))
            return torch.stack(m0, 0), torch.stack(m1, 0)
        keep = 1.0 - p
        m0 = torch.empty(steps, B, DH, dtype=torch.float32, device=device).bernoulli_(keep).div_(keep)
        m1 = torch.empty(steps, B, DO, dtype=torch.float32, device=device).bernoulli_(keep).div_(keep)
        return m0, m1

    def forward(self, enc_h, enc_l, dec_y=None, teacher_forcing_rate: float = 1, init_force: bool = False):
        B, T_enc, _ = enc_h.shape
        if self.training:
            steps = dec_y.size(-1)
        else:
            steps = self.CHR_MAX_STEPS
        K, V, lens_dev = self.attention.project_memory(enc_h, enc_l)
        use_gold = None
        if self.training:
            # one host coin per step t > 0, drawn exactly like the reference (src/models.py:356-357)
            use_gold = [False] * steps
            if _MASK_OVERRIDE['coins'] is not None:
                draws = [_MASK_OVERRIDE['coins'].pop(0) for _ in range(1, steps)]
            else:
                # ONE call: torch.rand(n) consumes the CPU generator exactly like n calls of torch.rand(1) (same values, same state
                # afterwards; checked by tests/test_cpu_abi.py) but costs 20 us instead of 1.5 ms of GPU-idle host time per batch
                draws = torch.rand(max(steps - 1, 0)).tolist()
            for t in range(1, steps):
                use_gold[t] = bool(draws[t - 1] <= teacher_forcing_rate)
        drop0, drop1 = self._decoder_masks(steps, B, enc_h.device)
        c0, c1 = self.lstms.lstms[0], self.lstms.lstms[1]
        params = (self.char_emb.weight, self.cls.bias, c0.weight_ih, c0.weight_hh, c0.bias_ih, c0.bias_hh,
                  c1.weight_ih, c1.weight_hh, c1.bias_ih, c1.bias_hh, self.attention.query_map.weight,
                  self.attention.query_map.bias, self.init_query)
        logits, att0, chars = LF.speller_loop(K, V, lens_dev, params, steps=steps, heads=self.att_heads,
                                              sos_idx=self.CHR_SOS_IDX, pad_idx=self.CHR_PAD_IDX, training=self.training,
                                              dec_y=dec_y if self.training else None, use_gold=use_gold, drop0=drop0, drop1=drop1,
                                              init_force=bool(init_force), defer_param_grads=True)
        self.last_chars = chars                         # (steps, B) greedy indices, device-side (extra, not in reference)
        # reference returns the attention map of sample 0 as a CPU tensor (heads, T_enc, steps+1) (:349,377,385):
        # one D2H at the end instead of one blocking copy per step
        att_wgts = LF.host_copy_lazy(att0.detach().permute(1, 2, 0))
        return logits, att_wgts


class ListenAttendSpell(nn.Module):
    """reference src/models.py:500-527"""

    def __init__(self, listener_configs: dict, speller_configs: dict):
        super().__init__()
        self.listener_configs = listener_configs
        self.speller_configs = speller_configs
        self.speller_configs['enc_out_dim'] = 2 * self.listener_configs['uniform_hid_dim']
        self.listen = Listener(**self.listener_configs)
        self.spell = Speller(**self.speller_configs)

    def forward(self, x, lx, dec_y=None, teacher_forcing_rate: float = 0.0, init_force: bool = False):
        enc_h, enc_l = self.listen(x, lx)
        pred_logits, att_wgts_list = self.spell(enc_h, enc_l, dec_y, teacher_forcing_rate, init_force)
        return pred_logits, att_wgts_list
