"""One LAS train step at the benchmarked shape (best config, B=96, T=1600, L=300, bf16 mode, yml dropouts, reducer -> backward overlap
eligible) between cudaProfilerStart / Stop, for `ncu --profile-from-start off` launch lists of exactly one step.
    python scripts/ncu_step.py [warmup_steps=2]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'attention-based-e2e-asr-dnn_b200'))
from las_b200 import configs as gu                      # noqa: E402
from las_b200.ddp import BucketedGradReducer            # noqa: E402
from las_b200.loss import masked_ce                     # noqa: E402
from las_b200.models import ListenAttendSpell           # noqa: E402
from las_b200.optim import FusedAdamW                   # noqa: E402

nw = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device('cuda:0')
B, T, L = 96, 1600, 300
cfg = gu.get_config('best', init_dropout=0.3, mid_dropout=0.3, final_dropout=0.35, dec_lstm_dropout=0.3)
torch.manual_seed(11785)
model = ListenAttendSpell(**cfg).to(dev).train()
opt = FusedAdamW(model.parameters(), lr=5e-4, weight_decay=5e-6, amsgrad=True)
red = BucketedGradReducer(list(model.named_parameters()), world_size=1)
x, lx, y = gu.make_inputs(11785, B, T, L)
x, y, lx = torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev), torch.from_numpy(lx)
ly = torch.full((B,), L, dtype=torch.int64)


def step():
    red.zero_grad()
    with torch.autocast('cuda', dtype=torch.bfloat16):
        logits, _ = model(x, lx, y, 1.0, False)
    loss, _ = masked_ce(logits, y, ly)
    (loss * 65536.0).backward()
    red.finish()
    opt.step_fused(inv_scale=1.0 / 65536.0, max_norm=5.0)
    return loss


for _ in range(nw):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
loss = step()
e1.record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(f'step ok: loss {float(loss):.4f}, {e0.elapsed_time(e1):.2f} ms (not a bench number when run under a profiler); '
      f'LAS_BWD_OVERLAP={os.environ.get("LAS_BWD_OVERLAP", "1")}')
