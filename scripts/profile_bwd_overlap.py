"""Kernel timeline of one train step's encoder backward (chrome trace via torch.profiler): which stream each BPTT kernel / GEMM ran
on and when, to check that the weight-gradient GEMMs of layer l+1 run beside the BPTT kernel of layer l (functional.py, backward
overlap).  Writes gpurun_out/bwd_overlap_trace.json and prints the BPTT kernels and the GEMMs that overlap them."""
import json, os, sys
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
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
path = os.path.join(ROOT, 'gpurun_out', 'bwd_overlap_trace.json')
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))['traceEvents'] if e.get('cat') in ('kernel', 'gpu_memcpy', 'gpu_memset') and 'dur' in e]
ev.sort(key=lambda e: e['ts'])
t0 = ev[0]['ts']
big = [e for e in ev if e['dur'] > 150]
for e in big:
    print(f"{(e['ts'] - t0) / 1e3:9.3f} ms  +{e['dur'] / 1e3:7.3f} ms  stream {e['args'].get('stream')}  grid {e['args'].get('grid')}  {e['name'][:70]}")
