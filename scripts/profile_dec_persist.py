"""Timeline of the persistent decoder-step kernel (csrc/decoder_persist.cu): %globaltimer stamps of attention CTA 0 and of the first
cell-0 / cell-1 CTA per decoder step.  Run on the GPU box:  python scripts/profile_dec_persist.py [B] [T_enc] [steps]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'attention-based-e2e-asr-dnn_b200'))
from las_b200 import _lib, configs                      # noqa: E402
from las_b200.models import ListenAttendSpell           # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 96
T = int(sys.argv[2]) if len(sys.argv) > 2 else 200
L = int(sys.argv[3]) if len(sys.argv) > 3 else 300
lib = C.CDLL(_lib.load()._name)
dev = torch.device('cuda:0')
torch.manual_seed(1)
model = ListenAttendSpell(**configs.get_config('best')).to(dev).train()
enc_h = torch.randn(B, T, 1024, device=dev) * 0.3
enc_l = torch.full((B,), T, dtype=torch.int64)
y = torch.randint(1, 29, (B, L), device=dev)
dbg = torch.zeros(256 * 16, dtype=torch.int64, device=dev)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for it in range(4):
    if it == 3:
        lib.las_dec_persist_set_debug(C.c_void_p(dbg.data_ptr()))
    ev0.record()
    with torch.autocast('cuda', dtype=torch.bfloat16):
        logits, _ = model.spell(enc_h, enc_l, y, 1.0, False)
    ev1.record()
    torch.cuda.synchronize()
    print(f'speller forward: {ev0.elapsed_time(ev1):.3f} ms for {L} steps = {ev0.elapsed_time(ev1) * 1e3 / L:.2f} us/step')
lib.las_dec_persist_set_debug(C.c_void_p(0))
d = dbg.cpu().numpy().reshape(256, 16).astype(np.float64)
lo, hi = 20, min(200, L - 2)
names = {0: 'ATT: c1 hand-off seen', 1: 'ATT: q projected', 14: 'ATT: K streamed (energies)', 2: 'ATT: V streamed (context)', 3: 'ATT: signalled',
         4: 'C0: own state loaded', 5: 'C0: att hand-off seen', 6: 'C0: ctx loaded', 7: 'C0: UMMAs retired', 9: 'C0: gates activated',
         11: 'C0: c / h stored', 15: 'C0: gates stored', 8: 'C0: signalled',
         10: 'C1: c0 hand-off seen', 12: 'C1: UMMAs retired', 13: 'C1: signalled'}
base = d[lo:hi, 0]
print(f'mean offsets from "ATT: c1 hand-off seen" of the same step, steps {lo}..{hi} (us):')
for k in (0, 1, 14, 2, 3, 4, 5, 6, 7, 9, 11, 15, 8, 10, 12, 13):
    print(f'  {names[k]:32s} {np.mean(d[lo:hi, k] - base) / 1e3:8.2f}')
print(f'step period: {np.mean(np.diff(d[lo:hi, 0])) / 1e3:.2f} us')
