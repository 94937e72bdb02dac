"""Times the fp32 (FFMA) GEMM launches that sit on the critical path of the AMP train step around the decoder loops
(csrc/decoder.cu): the tied classifier over all steps, dQC = dlogits . emb, and the batched dK / dV products.
    python scripts/bench_gemm_f32_shapes.py            # B=96, L=300, T_enc=200, P=256, V=30
CUDA events on the launching stream, 30 launches each after 5 warm-up launches; operands are re-used (L2-warm for the small ones,
the 59 MB outputs are not)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'attention-based-e2e-asr-dnn_b200'))
from las_b200 import functional as LF      # noqa: E402

dev = torch.device('cuda:0')
B, S, T, P, V = 96, 300, 200, 256, 30
E = 2 * P


def timed(fn, n=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / n


QC = torch.randn(S * B, 2 * P, device=dev)
emb = torch.randn(V, E, device=dev)
bias = torch.randn(V, device=dev)
logits = torch.empty(B, S, V, device=dev)
dlogits = torch.randn(B, S, V, device=dev)
dQC = torch.empty(S * B, 2 * P, device=dev)
DE = torch.randn(S + 1, B, T, device=dev)
Q = torch.randn(S + 1, B, 2 * P, device=dev)
dK = torch.empty(B, T, P, device=dev)

rows = []
# classifier: row m = t*B + b of QC -> logits[b, t, :]
rows.append(('classifier (28800 x 30, K = 512)', timed(lambda: LF.gemm_raw(QC, emb, logits, S * B, V, 2 * P, am=(0, 2 * P, 0), ak=(0, 1, 0),
                                                                            bk=(0, 1, 0), bn=E, cm=(V, S * V, B), bias1=bias)), 2.0 * S * B * V * 2 * P))
# dQC = dlogits . emb
rows.append(('dQC (28800 x 512, K = 30)', timed(lambda: LF.gemm_raw(dlogits, emb, dQC, S * B, 2 * P, V, am=(V, S * V, B), ak=(0, 1, 0),
                                                                     bk=(0, E, 0), bn=1, cm=(0, 2 * P, 0))), 2.0 * S * B * V * 2 * P))
# dK[b] = DE[:, b]^T . Q[:, b]
rows.append(('dK / dV (96 x (200 x 256), K = 301)', timed(lambda: LF.gemm_raw(DE, Q, dK, T, P, S + 1, am=(0, 1, 0), ak=(0, B * T, 0),
                                                                               bk=(0, B * 2 * P, 0), bn=1, cm=(0, P, 0), batch=B, bsA=T,
                                                                               bsB=2 * P, bsC=T * P)), 2.0 * B * T * P * (S + 1)))
for name, us, fl in rows:
    print(f'{name:42s} {us:8.1f} us  {fl / us / 1e6:6.2f} TFLOP/s')
