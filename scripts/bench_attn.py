"""Timing of the fused attention step kernel (fwd, bwd) at train (B=96,T=200) and greedy (B=256,T=375) shapes."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'attention-based-e2e-asr-dnn_b200'))
import torch
from las_b200 import functional as LF
DEV = 'cuda:0'
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=DEV)
for B, T, P in [(96, 200, 256), (256, 375, 256)]:
    q = torch.randn(B, P, device=DEV); K = torch.randn(B, T, P, device=DEV); V = torch.randn(B, T, P, device=DEV)
    lens = torch.full((B,), T, dtype=torch.int32, device=DEV)
    for mode in ('L2-warm (K/V re-read every decoder step)', 'cold (L2 flushed)'):
        ts = []
        for i in range(13):
            if mode.startswith('cold'):
                flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            NREP = 1 if mode.startswith('cold') else 20       # warm: 20 back-to-back launches amortise the host enqueue gap
            e0.record()
            for _ in range(NREP):
                LF.AttnStepFunction.apply(q, K, V, lens, 1)
            e1.record(); torch.cuda.synchronize()
            if i >= 3:
                ts.append(e0.elapsed_time(e1) / NREP)
        ts.sort(); t = ts[len(ts) // 2]
        byts = 2.0 * B * T * P * 4
        print(f'attn fwd B={B} T={T}: {mode:45s} {t*1e3:7.1f} us  {byts/t/1e6:8.1f} GB/s (algorithmic {byts/1e6:.1f} MB)')
