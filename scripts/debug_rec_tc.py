import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'attention-based-e2e-asr-dnn_b200'))
import numpy as np, torch
from las_b200 import _lib, functional as LF
lib = _lib.load()
DEV = 'cuda:0'
H, B, T = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
ndir, F = 2, 2 * H
rng = np.random.default_rng(0)
lens = [T] * B
if len(sys.argv) > 4:
    lens = [max(1, T - 2 * i) for i in range(B)]
if os.environ.get('LENS'):
    lens = [int(v) for v in os.environ['LENS'].split(',')]
print('lens', lens)
mask_t = torch.from_numpy(((rng.random((B, F)) > 0.3) / 0.7).astype(np.float32)).to(DEV) if len(sys.argv) > 5 else None
gates0 = torch.from_numpy(rng.standard_normal((B, T, ndir, 4 * H)).astype(np.float32)).to(DEV)
w_hh = torch.from_numpy(rng.uniform(-1, 1, size=(ndir, 4 * H, H)).astype(np.float32) / np.sqrt(H)).to(DEV)
if not os.environ.get('NOROUND'):
    w_hh = w_hh.to(torch.bfloat16).float()      # make weights exactly representable: isolates layout bugs from rounding
lens_dev = torch.tensor(lens, dtype=torch.int32, device=DEV)
res = []
for tc in (False, True):
    gates = gates0.clone()
    hs = torch.full((B, T + 2, F), 7.0, device=DEV); cs = torch.full((B, T + 2, F), 7.0, device=DEV); out = torch.full((B, T, F), 7.0, device=DEV)
    st = torch.cuda.current_stream().cuda_stream
    if tc:
        wb = LF.cast_bf16(w_hh, ndir * 4 * H, H, H, H)
        nbytes = lib.las_lstm_rec_tc_workspace_bytes(B, H, ndir); ws = torch.zeros(nbytes, dtype=torch.uint8, device=DEV)
        _lib.check(lib.las_lstm_rec_fwd_tc(gates.data_ptr(), wb.data_ptr(), lens_dev.data_ptr(), (mask_t.data_ptr() if mask_t is not None else None), out.data_ptr(), hs.data_ptr(), cs.data_ptr(), B, T, H, ndir, 1, ws.data_ptr(), nbytes, st), 'tc')
    else:
        nbytes = lib.las_lstm_rec_workspace_bytes(B, H, ndir); ws = torch.zeros(nbytes, dtype=torch.uint8, device=DEV)
        _lib.check(lib.las_lstm_rec_fwd_f32(gates.data_ptr(), w_hh.data_ptr(), lens_dev.data_ptr(), (mask_t.data_ptr() if mask_t is not None else None), out.data_ptr(), hs.data_ptr(), cs.data_ptr(), B, T, H, ndir, ws.data_ptr(), nbytes, st), 'f32')
    torch.cuda.synchronize()
    res.append((gates, hs, cs, out))
(g0, h0, c0, o0), (g1, h1, c1, o1) = res
for d in range(2):
    print('dir', d, 'per-t max |dh|:', [round(float((o1[:, t, d*H:(d+1)*H] - o0[:, t, d*H:(d+1)*H]).abs().max()), 4) for t in range(T)])
t = 1
dg = (g1[:, t, 0] - g0[:, t, 0]).abs()      # (B, 4H) dir 0, second step
print('step1 dir0 gate err by gate:', [round(float(dg[:, g*H:(g+1)*H].max()), 4) for g in range(4)])
print('step1 dir0 gate err by batch row:', [round(float(dg[b].max()), 3) for b in range(min(B, 8))])
print('step1 dir0 gate err by unit (gate0):', [round(float(dg[:, u].max()), 3) for u in range(0, min(H, 64), 4)])
# what does the tc kernel effectively compute? compare pre-activation delta
pre0 = torch.logit(g0[:, t, 0, :H].clamp(1e-6, 1-1e-6)) - gates0[:, t, 0, :H]
pre1 = torch.logit(g1[:, t, 0, :H].clamp(1e-6, 1-1e-6)) - gates0[:, t, 0, :H]
print('rec contribution (gate i) ref [b0,:6]:', pre0[0, :6].tolist())
print('rec contribution (gate i) tc  [b0,:6]:', pre1[0, :6].tolist())
hprev = h0[:, 1, :H]        # h at t=0, dir 0
ref = hprev @ w_hh[0, :H].t()
print('check ref formula [b0,:6]:', ref[0, :6].tolist())

for b in range(B):
    print('row', b, 'len', lens[b], 'max|dout| per t:', [round(float((o1[b, t] - o0[b, t]).abs().max()), 3) for t in range(T)])
