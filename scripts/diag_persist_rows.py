"""Where does the persistent decoder-step kernel's error against the fp32 oracle come from at B > 148 / on random encodings?
Prints max error, max |logit| and the same numbers for the launch-per-stage loop (LAS_DEC_PERSIST=0)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'attention-based-e2e-asr-dnn_b200'))
from helpers import gu, orc
from las_b200.models import ListenAttendSpell

def run(B, T_enc, L, scale_x, persist, seed=5):
    os.environ['LAS_DEC_PERSIST'] = persist
    cfg = gu.get_config('best'); sd = gu.make_state_dict(cfg, seed)
    model = ListenAttendSpell(**gu.get_config('best')).cuda().train()
    model.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd.items()})
    rng = np.random.default_rng(seed + 1)
    enc_h = torch.from_numpy((rng.standard_normal((B, T_enc, 1024)) * scale_x).astype(np.float32))
    y = torch.from_numpy(rng.integers(1, 29, size=(B, L)).astype(np.int64))
    lens = [T_enc, 1] + [int(v) for v in rng.integers(2, T_enc + 1, size=B - 2)]
    with torch.autocast('cuda', dtype=torch.bfloat16):
        logits, att = model.spell(enc_h.cuda(), torch.tensor(lens), y.cuda(), 1.0, False)
    p = {k: torch.from_numpy(v.copy()) for k, v in sd.items()}
    with torch.no_grad():
        ol, oatt = orc.speller_forward(p, enc_h, lens, heads=1, training=True, steps=L, dec_y=y, coins=[True] * L)
    e = np.abs(logits.detach().float().cpu().numpy() - ol.numpy())
    rows = e.max(axis=(1, 2))
    worst = np.argsort(rows)[-5:]
    print(f'B={B} x*{scale_x} persist={persist}: max err {e.max():.2e}  max|logit| {ol.abs().max():.2f}  mean err {e.mean():.2e} '
          f'per-step max {e.max(axis=(0, 2)).round(5).tolist()}  worst rows {worst.tolist()} lens {[lens[i] for i in worst]} '
          f'att err {np.abs(att.numpy() - oatt.numpy()).max():.2e}', flush=True)

for B, sx in ((5, 0.3), (5, 0.1), (200, 0.3), (200, 0.1), (148, 0.3), (160, 0.3)):
    for persist in ('1', '0'):
        run(B, 40, 5, sx, persist)
