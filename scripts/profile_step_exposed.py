"""What is on the train step's critical path besides the four latency-bound loops?  torch.profiler (CUPTI) timeline of one
step; "covered" = a recurrence kernel, the persistent decoder kernel or a kernel of the backward decoder loop is running.  The rest
of the step's wall time is attributed to the kernel that runs then (the earliest-started one when several overlap) or to 'idle'.
    python scripts/profile_step_exposed.py > gpurun_out/exposed.txt"""
import os, sys, collections
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
for _ in range(5): step()
torch.cuda.synchronize()
NS = 3
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(NS): step()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
LOOP = ('lstm_rec_', 'dec_persist', 'attn_bwd_tc_kernel', 'cell_bwd_kernel', 'gemm_bf16_tc_kernel<0, 1, 64', 'gemm_bf16_tc_kernel<false, true, 64', 'attn_step_split_kernel')
def is_loop(n): return any(k in n for k in LOOP)
t0, t1 = evs[0].time_range.start, max(e.time_range.end for e in evs)
# sweep: boundaries of all events
pts = sorted(set([e.time_range.start for e in evs] + [e.time_range.end for e in evs]))
import bisect
starts = [e.time_range.start for e in evs]
exposed = collections.Counter(); covered = 0.0
active = []
j = 0
for a, b in zip(pts[:-1], pts[1:]):
    while j < len(evs) and evs[j].time_range.start <= a:
        active.append(evs[j]); j += 1
    active = [e for e in active if e.time_range.end > a]
    if not active:
        exposed['(idle)'] += b - a
    elif any(is_loop(e.name) for e in active):
        covered += b - a
    else:
        exposed[min(active, key=lambda e: e.time_range.start).name[:90]] += b - a
print(f'{NS} steps: wall {(t1 - t0) / 1e3 / NS:.3f} ms per step; covered by the loops {covered / 1e3 / NS:.3f} ms; exposed {sum(exposed.values()) / 1e3 / NS:.3f} ms')
for n, g in exposed.most_common(40):
    print(f'  {g / 1e3 / NS:8.3f} ms  {n}')
# the same, in timeline order for one step (merged runs), to see where in the step the exposed pieces sit
print('\ntimeline of the LAST step (exposed runs >= 20 us):')
last0 = t0 + (t1 - t0) * (NS - 1) / NS
runs = []
active = []; j = 0
for a, b in zip(pts[:-1], pts[1:]):
    while j < len(evs) and evs[j].time_range.start <= a:
        active.append(evs[j]); j += 1
    active = [e for e in active if e.time_range.end > a]
    if a < last0: continue
    if not active: key = '(idle)'
    elif any(is_loop(e.name) for e in active): key = None
    else: key = min(active, key=lambda e: e.time_range.start).name[:70]
    if runs and runs[-1][0] == key and abs(runs[-1][2] - a) < 1e-6: runs[-1][2] = b
    else: runs.append([key, a, b])
# merge consecutive exposed runs into segments between covered spans
seg = []
for key, a, b in runs:
    if key is None:
        if seg:
            tot = sum(bb - aa for _, aa, bb in seg)
            if tot >= 20:
                c = collections.Counter()
                for k, aa, bb in seg: c[k] += bb - aa
                print(f'  at {(seg[0][1] - last0) / 1e3:7.3f} ms: {tot:7.1f} us exposed: ' + '; '.join(f'{k[:48]} {v:.0f}' for k, v in c.most_common(5)))
            seg = []
    else:
        seg.append((key, a, b))
