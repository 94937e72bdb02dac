"""GPU busy time vs wall time of one train step (torch.profiler / CUPTI): how much of the step is idle gaps between kernels."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'attention-based-e2e-asr-dnn_b200'))
import torch
from torch.profiler import profile, ProfilerActivity
from las_b200 import _lib, configs as gu
from las_b200.models import ListenAttendSpell
from las_b200.optim import FusedAdamW
from las_b200.ddp import BucketedGradReducer
from las_b200.loss import masked_ce
lib = _lib.load(); _lib.check(lib.las_init(0), 'init')
dev = torch.device('cuda:0')
B, T, L = 96, 1600, 300
cfg = gu.get_config('best'); torch.manual_seed(11785)
model = ListenAttendSpell(**cfg).to(dev).train()
opt = FusedAdamW(model.parameters(), lr=5e-4, weight_decay=5e-6, amsgrad=True)
red = BucketedGradReducer(list(model.named_parameters()), world_size=1)
x, lx, y = gu.make_inputs(1, B, T, L)
x, y, lx = torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev), torch.from_numpy(lx)
ly = torch.full((B,), L, dtype=torch.int64)
def step():
    red.zero_grad()
    with torch.autocast('cuda', dtype=torch.bfloat16):
        logits, _ = model(x, lx, y, 1.0, False)
    loss, _ = masked_ce(logits, y, ly)
    (loss * 65536.0).backward(); red.finish(); opt.step_fused(inv_scale=1.0 / 65536.0, max_norm=5.0)
for _ in range(4): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0, t1 = evs[0].time_range.start, max(e.time_range.end for e in evs)
busy, cur_end, gaps, prev = 0.0, t0, [], ''
for e in evs:
    s, en = e.time_range.start, e.time_range.end
    if s > cur_end:
        gaps.append((s - cur_end, e.name[:50] + '   <- after ' + prev[:50], cur_end, s)); busy += en - s; cur_end = en; prev = e.name
    elif en > cur_end:
        busy += en - cur_end; cur_end = en; prev = e.name
print(f'3 steps: wall {(t1 - t0) / 1e3:.2f} ms, GPU busy {busy / 1e3:.2f} ms, idle {(t1 - t0 - busy) / 1e3:.2f} ms in {len(gaps)} gaps')
gaps.sort(reverse=True)
cpu = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CPU]
for i, (g, n, g0, g1) in enumerate(gaps[:15]):
    print(f'  gap {g:8.1f} us before {n}')
    if i < 4:          # what the host was doing while the GPU sat idle
        for e in sorted(cpu, key=lambda e: e.time_range.start):
            a, b = max(e.time_range.start, g0), min(e.time_range.end, g1)
            if b - a > 40: print(f'        host {b - a:7.1f} us of {e.name[:70]} (whole call {e.time_range.end - e.time_range.start:.0f} us)')
import collections
small = collections.Counter(); 
for g, n, _, _ in gaps:
    small[n.split('<')[0][:40]] += g
for n, g in small.most_common(12): print(f'  total gap {g / 1e3:7.2f} ms before {n}')
# per-kernel device time (sum of durations; kernels on the side stream overlap others, so the sum exceeds the busy time)
tot = collections.Counter(); cnt = collections.Counter()
for e in evs:
    tot[e.name[:90]] += e.time_range.end - e.time_range.start; cnt[e.name[:90]] += 1
print('per step (3 steps averaged):')
for n, g in tot.most_common(40): print(f'  {g / 3e3:8.3f} ms  x{cnt[n] / 3:7.1f}  {n}')
