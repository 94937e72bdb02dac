"""Where does the bf16-mode logit error at the benchmarked lengths come from?  Runs the best_train_T1600_L300 fixture in bf16 mode with
parts of the tensor-pipe path switched back to fp32 kernels (one subprocess per variant, the switches are read at import / call time).
    python scripts/diag_bf16_error.py            # on the GPU box
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VARIANTS = {
    'default': {},
    'dec_fp32': {'LAS_DEC_TC': '0'},
    'rec_fp32': {'LAS_REC_TC': '0'},
    'dec+rec_fp32': {'LAS_DEC_TC': '0', 'LAS_REC_TC': '0'},
    'fuseq_off': {'LAS_DEC_FUSEQ': '0'},
}
CHILD = r'''
import sys, os, json, numpy as np, torch
sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'attention-based-e2e-asr-dnn_b200'))
import test_gpu_bench_shapes as T
from helpers import load_golden
g = load_golden('best_train_T1600_L300')
model, logits, att, loss = T._run_train_fixture(g, amp=True)
err = np.abs(logits - g['logits'])
wn, ws = T._grad_errors(model, g)
print('RESULT', json.dumps(dict(logits_abs=float(err.max()), logits_rms=float(np.sqrt((err**2).mean())), per_row=[float(err[b].max()) for b in range(err.shape[0])],
      argmax_t=[int(err[b].max(-1).argmax()) for b in range(err.shape[0])], gradnorm=wn, gradsample=ws)))
'''
for name, env in VARIANTS.items():
    e = dict(os.environ, **env)
    out = subprocess.run([sys.executable, '-c', 'ROOT=%r\n' % ROOT + CHILD], env=e, capture_output=True, text=True)
    line = [l for l in out.stdout.splitlines() if l.startswith('RESULT')]
    print(name, line[0][7:] if line else ('FAILED ' + out.stderr[-800:]), flush=True)
