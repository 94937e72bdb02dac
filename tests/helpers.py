"""Shared test helpers: run the oracle restatement on a golden fixture / on seeded inputs."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import golden_util as gu          # noqa: E402
from oracle import las_oracle as orc          # noqa: E402

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False)


def fixture_cfg(g):
    d = [float(v) for v in g['dropout']] if 'dropout' in g.files else [0, 0, 0, 0]
    over = {}
    if any(d):
        over = dict(init_dropout=d[0], mid_dropout=d[1], final_dropout=d[2], dec_lstm_dropout=d[3])
    if 'max_steps' in g.files:
        over['CHR_MAX_STEPS'] = int(g['max_steps'])
    return gu.get_config(str(g['cfg_name']), **over)


def fixture_masks(g, dtype=torch.float32):
    """(listener_masks, drop_masks) in the oracle's format, or (None, None)."""
    nl, nd = int(g['n_locked']), int(g['n_drops'])
    lm = [torch.from_numpy(g[f'locked_mask_{i}']).to(dtype) for i in range(nl)] if nl else None
    dm = None
    if nd:
        dm = [(torch.from_numpy(g[f'drop_mask_{2 * t}']).to(dtype), torch.from_numpy(g[f'drop_mask_{2 * t + 1}']).to(dtype))
              for t in range(nd // 2)]
    return lm, dm


def fixture_coins(g, steps):
    """coins[t] for t in 0..steps-1 (coins[0] unused): reference draws one torch.rand(1) per step t >= 1."""
    c = [False] * steps
    draws = g['coins']
    tf = float(g['tf_rate'])
    for t in range(1, steps):
        c[t] = bool(draws[t - 1] <= tf)
    return c


def oracle_train_from_fixture(g, dtype=torch.float32):
    cfg = fixture_cfg(g)
    sd = gu.make_state_dict(cfg, int(g['seed']))
    p = {k: torch.from_numpy(v.copy()).to(dtype).requires_grad_(True) for k, v in sd.items() if k != 'spell.cls.weight'}
    p['spell.cls.weight'] = p['spell.char_emb.weight']
    lc, sc = cfg['listener_configs'], cfg['speller_configs']
    y = torch.from_numpy(g['y'])
    L = y.shape[1]
    lm, dm = fixture_masks(g, dtype)
    logits, att = orc.las_forward(p, torch.from_numpy(g['x']).to(dtype), g['lx'].tolist(), lstm_layers=lc['lstm_layers'],
                                  plstm_layers=lc['plstm_layers'], heads=sc['att_heads'], training=True, steps=L,
                                  dec_y=y, coins=fixture_coins(g, L), listener_masks=lm, drop_masks=dm,
                                  init_force=bool(g['init_force']))
    loss = orc.masked_ce_loss(logits, y, g['ly'].tolist())
    loss.backward()
    grads = {k: v.grad for k, v in p.items() if k != 'spell.cls.weight'}
    return logits.detach(), att, loss.detach(), grads


def rel_err(a, b, floor=1e-30):
    """max|a-b| / max(max|b|, floor).  `floor` guards tensors that are mathematically zero (e.g. the gradient of
    spell.attention.key_map.bias: a constant added to every energy of a row cancels in the softmax)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), floor))


def grad_floor(g):
    """1e-3 x the largest gradient norm in a training fixture: the scale below which a gradient is round-off."""
    return 1e-3 * max(float(g[k]) for k in g.files if k.startswith('gradnorm.'))
