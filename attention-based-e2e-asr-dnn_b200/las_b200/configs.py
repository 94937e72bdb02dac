"""Model configurations of the reference's YAML schema (config/sample-attention.yml) as Python dicts, and seeded synthetic
inputs of BASELINE.json's shapes -- what bench.py and examples feed the product path (no dependency on oracle/).

    model = ListenAttendSpell(**get_config('best'))
"""
from __future__ import annotations

import copy
from typing import Dict, List, Tuple

import numpy as np

# speller_configs carries dec_vocab_size / CHR_SOS_IDX / CHR_PAD_IDX the way src/train.py:503-505 injects them
CONFIGS: Dict[str, dict] = {
    # BASELINE.json configs[0]: tiny base-LAS (hid 128, 1 pLSTM layer)
    'tiny': dict(
        listener_configs=dict(input_dim=15, uniform_hid_dim=128, lstm_layers=1, plstm_layers=1, bidirectional=True,
                              init_dropout=0.0, mid_dropout=0.0, final_dropout=0.0),
        speller_configs=dict(att_proj_dim=64, att_heads=1, att_dropout=0.0, dec_emb_dim=128, dec_emb_dropout=0.0,
                             dec_lstm_hid_dim=128, dec_lstm_out_dim=64, dec_lstm_dropout=0.0, CHR_MAX_STEPS=40,
                             USE_GREEDY=True, dec_vocab_size=30, CHR_SOS_IDX=0, CHR_PAD_IDX=29)),
    # config/sample-attention.yml:42-68, the "best" base-LAS of BASELINE.json configs[1..3] (dropouts left to the caller)
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


def make_inputs(seed: int, B: int, T: int, L: int, lx: List[int] = None, input_dim: int = 15
                ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Synthetic batch of SURVEY.md 8(d): x ~ N(0,1) MFCC-like (B, T, input_dim) with the pad region zeroed like pad_sequence,
    lx (B,) int64 (all T unless given), dec_y ~ randint(1, 29) (B, L) int64 (letters / apostrophe / space only)."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal(size=(B, T, input_dim)).astype(np.float32)
    lx = np.full((B,), T, dtype=np.int64) if lx is None else np.asarray(lx, dtype=np.int64)
    for b in range(B):
        x[b, lx[b]:] = 0.0
    y = rng.integers(1, 29, size=(B, L)).astype(np.int64)
    return x, lx, y


# config/rewriter.yml:46-63 -- the char-to-char attention seq2seq LM of src/lmtrain.py (BASELINE.json configs[4])
REWRITER_CONFIGS: Dict[str, dict] = {
    'rw_yml': dict(vocab_size=30, emb_dim=256, enc_lstm_layers=2, enc_lstm_hid_dim=256, enc_dropouts=[0.3, 0.3], att_proj_dim=128,
                   att_heads=4, att_dropout=0.2, dec_lstm_layers=2, dec_lstm_hid_dim=256, dec_lstm_out_dim=128, dec_lstm_dropout=0.3,
                   CHR_PAD_IDX=29, CHR_MAX_STEPS=600, CHR_SOS_IDX=0),
}


def get_rewriter_config(name: str = 'rw_yml', **overrides) -> dict:
    cfg = copy.deepcopy(REWRITER_CONFIGS[name])
    cfg.update(overrides)
    return cfg


def make_token_inputs(seed: int, B: int, Tx: int, L: int, lx=None):
    """Rewriter inputs: x tokens (B, Tx) in 1..28 padded with 29 past each length, lx, y tokens (B, L)."""
    rng = np.random.default_rng(seed)
    lx = np.asarray(lx if lx is not None else [Tx] * B, dtype=np.int64)
    x = rng.integers(1, 29, size=(B, Tx)).astype(np.int64)
    for b in range(B):
        x[b, lx[b]:] = 29
    y = rng.integers(1, 29, size=(B, L)).astype(np.int64)
    return x, lx, y
