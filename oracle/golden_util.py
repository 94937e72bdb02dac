"""
ORACLE -- TEST INFRASTRUCTURE ONLY.

Deterministic (numpy, platform-stable) builders for LAS configurations, weights and synthetic inputs,
shared by oracle/make_golden.py (which feeds them to the unmodified reference) and by the tests (which feed
the same weights/inputs to the oracle restatement and to the CUDA path).  Nothing here computes the model.
"""
from __future__ import annotations

import copy
from typing import Dict, List, Tuple

import numpy as np

# name -> (listener_configs, speller_configs); speller gets dec_vocab_size / CHR_SOS_IDX / CHR_PAD_IDX the way
# src/train.py:503-505 injects them.
CONFIGS: Dict[str, dict] = {
    # smallest shape that still exercises 1 LSTM + 3 pLSTM levels
    'micro': dict(
        listener_configs=dict(input_dim=15, uniform_hid_dim=32, lstm_layers=1, plstm_layers=3, bidirectional=True,
                              init_dropout=0.0, mid_dropout=0.0, final_dropout=0.0),
        speller_configs=dict(att_proj_dim=16, att_heads=1, att_dropout=0.0, dec_emb_dim=32, dec_emb_dropout=0.0,
                             dec_lstm_hid_dim=32, dec_lstm_out_dim=16, dec_lstm_dropout=0.0, CHR_MAX_STEPS=12,
                             USE_GREEDY=True, dec_vocab_size=30, CHR_SOS_IDX=0, CHR_PAD_IDX=29)),
    # BASELINE.json configs[0] / SURVEY.md Appendix B "legal tiny config"
    'tiny': dict(
        listener_configs=dict(input_dim=15, uniform_hid_dim=128, lstm_layers=1, plstm_layers=1, bidirectional=True,
                              init_dropout=0.0, mid_dropout=0.0, final_dropout=0.0),
        speller_configs=dict(att_proj_dim=64, att_heads=1, att_dropout=0.0, dec_emb_dim=128, dec_emb_dropout=0.0,
                             dec_lstm_hid_dim=128, dec_lstm_out_dim=64, dec_lstm_dropout=0.0, CHR_MAX_STEPS=40,
                             USE_GREEDY=True, dec_vocab_size=30, CHR_SOS_IDX=0, CHR_PAD_IDX=29)),
    # config/sample-attention.yml:42-68 (the "best" base-LAS), dropout left to the caller
    'best': dict(
        listener_configs=dict(input_dim=15, uniform_hid_dim=512, lstm_layers=1, plstm_layers=3, bidirectional=True,
                              init_dropout=0.0, mid_dropout=0.0, final_dropout=0.0),
        speller_configs=dict(att_proj_dim=256, att_heads=1, att_dropout=0.0, dec_emb_dim=512, dec_emb_dropout=0.0,
                             dec_lstm_hid_dim=512, dec_lstm_out_dim=256, dec_lstm_dropout=0.0, CHR_MAX_STEPS=600,
                             USE_GREEDY=True, dec_vocab_size=30, CHR_SOS_IDX=0, CHR_PAD_IDX=29)),
}


def get_config(name: str, **overrides) -> dict:
    cfg = copy.deepcopy(CONFIGS[name])
    for k, v in overrides.items():
        if k in cfg['listener_configs']:
            cfg['listener_configs'][k] = v
        elif k in cfg['speller_configs']:
            cfg['speller_configs'][k] = v
        else:
            raise KeyError(k)
    return cfg


def state_dict_shapes(cfg: dict) -> List[Tuple[str, Tuple[int, ...]]]:
    """The frozen state_dict contract, in named_parameters() order."""
    lc, sc = cfg['listener_configs'], cfg['speller_configs']
    H, nd = lc['uniform_hid_dim'], 2 if lc['bidirectional'] else 1
    out: List[Tuple[str, Tuple[int, ...]]] = []

    def lstm(prefix, din):
        for suf in ([''] if nd == 1 else ['', '_reverse']):
            out.append((f'{prefix}weight_ih_l0{suf}', (4 * H, din)))
            out.append((f'{prefix}weight_hh_l0{suf}', (4 * H, H)))
            out.append((f'{prefix}bias_ih_l0{suf}', (4 * H,)))
            out.append((f'{prefix}bias_hh_l0{suf}', (4 * H,)))

    for i in range(lc['lstm_layers']):
        lstm(f'listen.base.lstms.{i}.', lc['input_dim'] if i == 0 else nd * H)
    for i in range(lc['plstm_layers']):
        lstm(f'listen.pyramid.plstms.{i}.', 2 * nd * H)
    P, E, DH, DO, V = (sc['att_proj_dim'], sc['dec_emb_dim'], sc['dec_lstm_hid_dim'], sc['dec_lstm_out_dim'],
                       sc['dec_vocab_size'])
    enc = nd * H if nd == 2 else 2 * H   # ListenAttendSpell sets enc_out_dim = 2*uniform_hid_dim (models.py:512)
    out.append(('spell.init_query', (1, DO)))
    for nm, (o, i) in (('key_map', (P, enc)), ('value_map', (P, enc)), ('query_map', (P, DO)), ('final_map', (P, P))):
        out.append((f'spell.attention.{nm}.weight', (o, i)))
        out.append((f'spell.attention.{nm}.bias', (o,)))
    out.append(('spell.char_emb.weight', (V, E)))
    out.append(('spell.lstms.lstms.0.weight_ih', (4 * DH, E + P)))
    out.append(('spell.lstms.lstms.0.weight_hh', (4 * DH, DH)))
    out.append(('spell.lstms.lstms.0.bias_ih', (4 * DH,)))
    out.append(('spell.lstms.lstms.0.bias_hh', (4 * DH,)))
    out.append(('spell.lstms.lstms.1.weight_ih', (4 * DO, DH)))
    out.append(('spell.lstms.lstms.1.weight_hh', (4 * DO, DO)))
    out.append(('spell.lstms.lstms.1.bias_ih', (4 * DO,)))
    out.append(('spell.lstms.lstms.1.bias_hh', (4 * DO,)))
    out.append(('spell.cls.bias', (V,)))
    return out


def make_state_dict(cfg: dict, seed: int, scale: float = 1.0) -> Dict[str, np.ndarray]:
    """Seeded numpy weights with torch-default-like magnitudes.  `spell.cls.weight` aliases `spell.char_emb.weight`
    (src/models.py:287).  Row CHR_PAD_IDX of the embedding is zero like nn.Embedding(padding_idx=29) at init."""
    rng = np.random.default_rng(seed)
    sd: Dict[str, np.ndarray] = {}
    for name, shape in state_dict_shapes(cfg):
        if name == 'spell.init_query':
            w = rng.uniform(0.0, 1.0, size=shape)
        elif name == 'spell.char_emb.weight':
            w = rng.standard_normal(size=shape) * 0.5
            w[cfg['speller_configs']['CHR_PAD_IDX']] = 0.0
        else:
            fan = shape[-1] if len(shape) > 1 else None
            if 'lstm' in name:        # nn.LSTM / nn.LSTMCell: U(-1/sqrt(hidden), 1/sqrt(hidden))
                hid = shape[0] // 4
                k = 1.0 / np.sqrt(hid)
            elif fan is not None:     # nn.Linear weight
                k = 1.0 / np.sqrt(fan)
            else:                     # nn.Linear bias
                k = 0.05
            w = rng.uniform(-k, k, size=shape) * scale
        sd[name] = w.astype(np.float32)
    sd['spell.cls.weight'] = sd['spell.char_emb.weight']
    return sd


def make_inputs(seed: int, B: int, T: int, L: int, lx: List[int] = None, input_dim: int = 15
                ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """x ~ N(0,1) MFCC-like, pad region zeroed like pad_sequence (src/utils.py:114-116);
    dec_y ~ randint(1, 29) (letters / apostrophe / space only)."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal(size=(B, T, input_dim)).astype(np.float32)
    lx = np.full((B,), T, dtype=np.int64) if lx is None else np.asarray(lx, dtype=np.int64)
    for b in range(B):
        x[b, lx[b]:] = 0.0
    y = rng.integers(1, 29, size=(B, L)).astype(np.int64)
    return x, lx, y


# ---- Rewriter (src/lmtrain.py:95-253, config/rewriter.yml:46-63) ----
REWRITER_CONFIGS = {
    # fp32-kernel-sized fixture
    'rw_micro': dict(vocab_size=30, emb_dim=32, enc_lstm_layers=2, enc_lstm_hid_dim=32, enc_dropouts=[0.0, 0.0], att_proj_dim=16,
                     att_heads=4, att_dropout=0.0, dec_lstm_layers=2, dec_lstm_hid_dim=24, dec_lstm_out_dim=8, dec_lstm_dropout=0.0,
                     CHR_PAD_IDX=29, CHR_MAX_STEPS=12, CHR_SOS_IDX=0),
    # config/rewriter.yml dims
    'rw_yml': dict(vocab_size=30, emb_dim=256, enc_lstm_layers=2, enc_lstm_hid_dim=256, enc_dropouts=[0.3, 0.3], att_proj_dim=128,
                   att_heads=4, att_dropout=0.2, dec_lstm_layers=2, dec_lstm_hid_dim=256, dec_lstm_out_dim=128, dec_lstm_dropout=0.3,
                   CHR_PAD_IDX=29, CHR_MAX_STEPS=600, CHR_SOS_IDX=0),
}


def get_rewriter_config(name: str, **overrides) -> dict:
    cfg = copy.deepcopy(REWRITER_CONFIGS[name])
    cfg.update(overrides)
    return cfg


def rewriter_state_dict_shapes(cfg: dict) -> List[Tuple[str, Tuple[int, ...]]]:
    """Rewriter.state_dict() keys / shapes in named_parameters() order (verified against the reference in make_golden)."""
    V, E, H, P, DH, DO = (cfg['vocab_size'], cfg['emb_dim'], cfg['enc_lstm_hid_dim'], cfg['att_proj_dim'], cfg['dec_lstm_hid_dim'],
                          cfg['dec_lstm_out_dim'])
    out: List[Tuple[str, Tuple[int, ...]]] = [('init_query', (1, DO)), ('char_emb.weight', (V, E))]
    for i in range(cfg['enc_lstm_layers']):
        din = E if i == 0 else 2 * H
        for suf in ['', '_reverse']:
            out += [(f'enc_lstm.lstms.{i}.weight_ih_l0{suf}', (4 * H, din)), (f'enc_lstm.lstms.{i}.weight_hh_l0{suf}', (4 * H, H)),
                    (f'enc_lstm.lstms.{i}.bias_ih_l0{suf}', (4 * H,)), (f'enc_lstm.lstms.{i}.bias_hh_l0{suf}', (4 * H,))]
    for nm, (o, i) in (('key_map', (P, 2 * H)), ('value_map', (P, 2 * H)), ('query_map', (P, DO)), ('final_map', (P, P))):
        out += [(f'mha.{nm}.weight', (o, i)), (f'mha.{nm}.bias', (o,))]
    out += [('dec_lstm.lstms.0.weight_ih', (4 * DH, E + P)), ('dec_lstm.lstms.0.weight_hh', (4 * DH, DH)),
            ('dec_lstm.lstms.0.bias_ih', (4 * DH,)), ('dec_lstm.lstms.0.bias_hh', (4 * DH,)),
            ('dec_lstm.lstms.1.weight_ih', (4 * DO, DH)), ('dec_lstm.lstms.1.weight_hh', (4 * DO, DO)),
            ('dec_lstm.lstms.1.bias_ih', (4 * DO,)), ('dec_lstm.lstms.1.bias_hh', (4 * DO,)), ('cls.bias', (V,))]
    return out


def make_rewriter_state_dict(cfg: dict, seed: int, scale: float = 1.0) -> Dict[str, np.ndarray]:
    rng = np.random.default_rng(seed)
    sd = {}
    for k, shp in rewriter_state_dict_shapes(cfg):
        fan = shp[-1] if len(shp) > 1 else shp[0]
        bound = scale / np.sqrt(max(fan, 1))
        if k == 'init_query':
            sd[k] = rng.uniform(0, 1, size=shp).astype(np.float32)
        else:
            sd[k] = rng.uniform(-bound, bound, size=shp).astype(np.float32)
    sd['cls.weight'] = sd['char_emb.weight']
    return sd


def make_token_inputs(seed: int, B: int, Tx: int, L: int, lx=None):
    """Rewriter inputs: x tokens (B,Tx) in 1..28 padded with 29 past each length, lx, y tokens (B,L)."""
    rng = np.random.default_rng(seed)
    lx = np.asarray(lx if lx is not None else [Tx] * B, dtype=np.int64)
    x = rng.integers(1, 29, size=(B, Tx)).astype(np.int64)
    for b in range(B):
        x[b, lx[b]:] = 29
    y = rng.integers(1, 29, size=(B, L)).astype(np.int64)
    return x, lx, y
