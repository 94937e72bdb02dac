"""Rewriter -- the reference's char-to-char attention seq2seq LM (src/lmtrain.py:95-253) on the same kernels.

Same class name, constructor kwargs, forward signature, parameter names / shapes (state_dict keys: char_emb, enc_lstm,
mha, dec_lstm, cls, init_query) as the reference.  The encoder is the reference's LockedLSTM stack (no pyramid), the decoder
the same two-cell loop with projected dot-product attention as the Speller, run as one C call per direction
(las_speller_fwd_f32 / las_speller_bwd_f32).

Reference quirk preserved: the teacher-forcing branch assigns the gold embedding to a misspelt
variable (`char_meb`, src/lmtrain.py:231), so teacher forcing NEVER takes effect -- the decoder always feeds back its own
argmax -- but the coin `torch.rand(1)` is still drawn once per step t > 0 in training mode (RNG stream parity).
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from . import functional as LF
from .models import MultiheadCrossAttention
from .modules import AutoRegDecoderLSTMCell, LockedLSTM, _MASK_OVERRIDE


class Rewriter(nn.Module):
    def __init__(self, vocab_size: int = 30, emb_dim: int = 256, enc_lstm_layers: int = 3, enc_lstm_hid_dim: int = 256,
                 enc_dropouts: list = [0.3, 0.3], att_proj_dim: int = 128, att_heads: int = 4, att_dropout: float = 0.2,
                 dec_lstm_layers: int = 2, dec_lstm_hid_dim: int = 256, dec_lstm_out_dim: int = 128, dec_lstm_dropout: float = 0.3,
                 CHR_PAD_IDX: int = 29, CHR_MAX_STEPS: int = 600, CHR_SOS_IDX: int = 0):
        super().__init__()
        self.vocab_size = vocab_size
        self.emb_dim = emb_dim
        self.enc_lstm_layers = enc_lstm_layers
        self.enc_lstm_hid_dim = enc_lstm_hid_dim
        self.enc_dropouts = enc_dropouts
        self.att_proj_dim = att_proj_dim
        self.att_heads = att_heads
        self.att_dropout = att_dropout
        self.dec_lstm_layers = dec_lstm_layers
        self.dec_lstm_hid_dim = dec_lstm_hid_dim
        self.dec_lstm_out_dim = dec_lstm_out_dim
        self.dec_lstm_dropout = dec_lstm_dropout
        self.CHR_PAD_IDX = CHR_PAD_IDX
        self.CHR_MAX_STEPS = CHR_MAX_STEPS
        self.CHR_SOS_IDX = CHR_SOS_IDX
        if self.emb_dim != 2 * self.att_proj_dim:
            raise ValueError(f'emb_dim ({emb_dim}) must equal 2*att_proj_dim ({2 * att_proj_dim}): the tied classifier consumes '
                             'cat[q_proj, context] (reference src/lmtrain.py:174)')
        self.char_emb = nn.Embedding(num_embeddings=self.vocab_size, embedding_dim=self.emb_dim, padding_idx=self.CHR_PAD_IDX)
        self.enc_lstm = LockedLSTM(lstm_input_dim=self.emb_dim, uniform_hid_dim=self.enc_lstm_hid_dim, lstm_layers=self.enc_lstm_layers,
                                   bidirectional=True, init_dropout=self.enc_dropouts[0], mid_dropout=self.enc_dropouts[-1])
        self.mha = MultiheadCrossAttention(enc_out_dim=self.enc_lstm_hid_dim * 2, dec_out_dim=self.dec_lstm_out_dim,
                                           proj_dim=self.att_proj_dim, heads=self.att_heads, dropout=self.att_dropout)
        self.dec_lstm = AutoRegDecoderLSTMCell(att_proj_dim=self.att_proj_dim, dec_emb_dim=self.emb_dim, dec_hid_dim=self.dec_lstm_hid_dim,
                                               dec_out_dim=self.dec_lstm_out_dim, dec_mid_dropout=self.dec_lstm_dropout)
        self.cls = nn.Linear(self.emb_dim, self.vocab_size)
        self.cls.weight = self.char_emb.weight            # weight tying (src/lmtrain.py:177)
        self.init_query = nn.Parameter(torch.rand((1, self.dec_lstm_out_dim)), requires_grad=True)
        # unregistered, never trained, always zero (src/lmtrain.py:181-187) -- kept for attribute parity
        self.init_hiddens = [(nn.Parameter(torch.zeros((1, self.dec_lstm_hid_dim)), requires_grad=True),
                              nn.Parameter(torch.zeros((1, self.dec_lstm_hid_dim)), requires_grad=True)),
                             (nn.Parameter(torch.zeros((1, self.dec_lstm_out_dim)), requires_grad=True),
                              nn.Parameter(torch.zeros((1, self.dec_lstm_out_dim)), requires_grad=True))]

    def _decoder_masks(self, steps, B, device):
        p = self.dec_lstm.dec_mid_dropout
        if (not self.training) or (not p):
            return None, None
        DH, DO = self.dec_lstm_hid_dim, self.dec_lstm_out_dim
        if _MASK_OVERRIDE['drop'] is not None or os.environ.get('LAS_EXACT_RNG', '0') == '1':
            m0, m1 = [], []
            for _ in range(steps):
                m0.append(self.dec_lstm.draw_dropout_mask(B, DH, device))
                m1.append(self.dec_lstm.draw_dropout_mask(B, DO, device))
            return torch.stack(m0, 0), torch.stack(m1, 0)
        keep = 1.0 - p
        m0 = torch.empty(steps, B, DH, dtype=torch.float32, device=device).bernoulli_(keep).div_(keep)
        m1 = torch.empty(steps, B, DO, dtype=torch.float32, device=device).bernoulli_(keep).div_(keep)
        return m0, m1

    def forward(self, x, lx, dec_y=None, tf_rate: float = 1.0, init_force=None):
        # token embedding of the encoder input (src/lmtrain.py:191): a gather -- torch plumbing; everything below is ours
        xe = self.char_emb(x)
        enc_h, enc_l = self.enc_lstm(xe, lx)
        B = enc_h.shape[0]
        if self.training:
            steps = dec_y.size(-1)
        else:
            steps = self.CHR_MAX_STEPS
        K, V, lens_dev = self.mha.project_memory(enc_h, enc_l)
        use_gold = None
        if self.training:
            # the coin is drawn like the reference (src/lmtrain.py:229-230) but its outcome is never used (:231 assigns `char_meb`)
            use_gold = [False] * steps
            if _MASK_OVERRIDE['coins'] is not None:
                for t in range(1, steps):
                    _MASK_OVERRIDE['coins'].pop(0)
            else:
                torch.rand(max(steps - 1, 0))          # same generator consumption as steps-1 calls of torch.rand(1)
        drop0, drop1 = self._decoder_masks(steps, B, enc_h.device)
        c0, c1 = self.dec_lstm.lstms[0], self.dec_lstm.lstms[1]
        params = (self.char_emb.weight, self.cls.bias, c0.weight_ih, c0.weight_hh, c0.bias_ih, c0.bias_hh,
                  c1.weight_ih, c1.weight_hh, c1.bias_ih, c1.bias_hh, self.mha.query_map.weight, self.mha.query_map.bias,
                  self.init_query)
        logits, att0, chars = LF.speller_loop(K, V, lens_dev, params, steps=steps, heads=self.att_heads, sos_idx=self.CHR_SOS_IDX,
                                              pad_idx=self.CHR_PAD_IDX, training=self.training,
                                              dec_y=dec_y if self.training else None, use_gold=use_gold, drop0=drop0, drop1=drop1)
        self.last_chars = chars
        att_wgts = LF.host_copy_lazy(att0.detach().permute(1, 2, 0))       # (heads, T_enc, steps+1) CPU tensor like the reference (:249-251)
        return logits, att_wgts
