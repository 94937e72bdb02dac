"""Speller (decoder loop) alone at the train shape: B=96, T_enc=200, L=300, AMP mode.  Prints CUDA-event times for fwd / bwd and
is the target of `ncu --metrics gpu__time_duration.sum` for per-kernel durations."""
import sys, os, time
ROOT = os.path.join(os.path.dirname(__file__), '..')
sys.path.insert(0, os.path.join(ROOT, 'attention-based-e2e-asr-dnn_b200')); sys.path.insert(0, ROOT)
import torch
from las_b200.models import Speller
from oracle import golden_util as gu
DEV = 'cuda:0'
B, T, L = int(os.environ.get('B', 96)), int(os.environ.get('TENC', 200)), int(os.environ.get('L', 300))
cfg = gu.get_config('best')['speller_configs']; cfg['enc_out_dim'] = 1024
torch.manual_seed(0)
sp = Speller(**cfg).to(DEV).train()
enc_h = torch.randn(B, T, 1024, device=DEV, requires_grad=True)
enc_l = torch.full((B,), T, dtype=torch.int64)
y = torch.randint(1, 29, (B, L), device=DEV)
reps = int(os.environ.get('REPS', 3))
for it in range(reps):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    t0 = time.perf_counter()
    e[0].record()
    with torch.autocast('cuda', dtype=torch.bfloat16):
        logits, _ = sp(enc_h, enc_l, y, 1.0, False)
    t1 = time.perf_counter()
    e[1].record()
    logits.float().sum().backward()
    t2 = time.perf_counter()
    e[2].record(); torch.cuda.synchronize()
    print(f'iter {it}: fwd {e[0].elapsed_time(e[1]):.2f} ms (host enqueue {1e3*(t1-t0):.2f}), bwd {e[1].elapsed_time(e[2]):.2f} ms (host enqueue {1e3*(t2-t1):.2f})')
