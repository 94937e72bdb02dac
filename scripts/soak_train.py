"""Soak run: N train steps with a different ragged batch every step (random T, lengths, transcript lengths) through the shipped
schedule (reducer -> backward overlap, GEMM tiles behind the recurrence kernels' progress counters, persistent decoder kernel).
Checks that nothing hangs, the loss stays finite and the allocator's reserved memory settles.   python scripts/soak_train.py [steps=200]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'attention-based-e2e-asr-dnn_b200'))
from las_b200 import configs as gu                      # noqa: E402
from las_b200.ddp import BucketedGradReducer            # noqa: E402
from las_b200.loss import masked_ce                     # noqa: E402
from las_b200.models import ListenAttendSpell           # noqa: E402
from las_b200.optim import FusedAdamW                   # noqa: E402
from las_b200 import functional as LF                   # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 200
dev = torch.device('cuda:0')
cfg = gu.get_config('best', init_dropout=0.3, mid_dropout=0.3, final_dropout=0.35, dec_lstm_dropout=0.3)
torch.manual_seed(1)
model = ListenAttendSpell(**cfg).to(dev).train()
opt = FusedAdamW(model.parameters(), lr=1e-4, weight_decay=5e-6, amsgrad=True)
red = BucketedGradReducer(list(model.named_parameters()), world_size=1)
rng = np.random.default_rng(7)
t0 = time.time()
peak = []
for it in range(N):
    B = int(rng.integers(40, 97))
    T = int(rng.integers(60, 201)) * 8
    L = int(rng.integers(20, 301))
    lx = np.sort(rng.integers(max(8, T // 3), T + 1, size=B))[::-1].copy()
    lx[0] = T
    x, lxa, y = gu.make_inputs(100 + it, B, T, L, lx=list(lx))
    ly = torch.from_numpy(rng.integers(max(2, L // 2), L + 1, size=B).astype(np.int64))
    ly[0] = L
    xd, yd = torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev)
    tf = 1.0 if it % 3 else 0.7
    red.zero_grad()
    with torch.autocast('cuda', dtype=torch.bfloat16):
        logits, _ = model(xd, torch.from_numpy(lxa), yd, tf, False)
    loss, _ = masked_ce(logits, yd, ly)
    (loss * 1024.0).backward()
    red.finish()
    opt.step_fused(inv_scale=1.0 / 1024.0, max_norm=5.0)
    if it % 20 == 19 or it == N - 1:
        lv = float(loss.detach())
        peak.append(torch.cuda.memory_reserved() / 2 ** 30)
        print(f'step {it + 1}: B={B} T={T} L={L} tf={tf} loss {lv:.4f} reserved {peak[-1]:.2f} GiB, pool {LF.speller_pool_bytes() / 2 ** 20:.0f} MiB, '
              f'{(time.time() - t0):.1f} s', flush=True)
        assert np.isfinite(lv), 'loss is not finite'
torch.cuda.synchronize()
assert peak[-1] <= 1.05 * max(peak[len(peak) // 2:]) + 0.01, peak
print(f'soak ok: {N} ragged steps, last reserved {peak[-1]:.2f} GiB')
