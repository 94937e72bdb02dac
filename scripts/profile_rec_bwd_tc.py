"""Per-phase timeline of the tensor-pipe BPTT kernel (clock64 stamps of CTA (0,0,0))."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'attention-based-e2e-asr-dnn_b200'))
import numpy as np, torch
from las_b200 import _lib, functional as LF
lib = _lib.load()
DEV = 'cuda:0'
H, B, T = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
ndir, F = 2, 2 * H
rng = np.random.default_rng(0)
st = torch.cuda.current_stream().cuda_stream
gates = torch.from_numpy(rng.standard_normal((B, T, ndir, 4 * H)).astype(np.float32)).to(DEV)
w_hh = torch.from_numpy((rng.uniform(-1, 1, size=(ndir, 4 * H, H)) / np.sqrt(H)).astype(np.float32)).to(DEV)
lens_dev = torch.full((B,), T, dtype=torch.int32, device=DEV)
hs = torch.zeros(B, T + 2, F, device=DEV); cs = torch.zeros(B, T + 2, F, device=DEV)
dout = torch.randn(B, T, F, device=DEV)
wb = LF.cast_bf16(w_hh, ndir * 4 * H, H, H, H)
nbytes = lib.las_lstm_rec_tc_workspace_bytes(B, H, ndir); ws = torch.zeros(nbytes, dtype=torch.uint8, device=DEV)
_lib.check(lib.las_lstm_rec_fwd_tc(gates.data_ptr(), wb.data_ptr(), lens_dev.data_ptr(), None, None, hs.data_ptr(), cs.data_ptr(), B, T, H, ndir, 1, ws.data_ptr(), nbytes, st), 'fwd')
w_t = torch.empty(ndir, H, 4 * H, dtype=torch.bfloat16, device=DEV)
_lib.check(lib.las_transpose_cast_bf16(w_hh.data_ptr(), w_t.data_ptr(), ndir, 4 * H, H, st), 'tr')
dgb = torch.empty(B * T, ndir * 4 * H, dtype=torch.bfloat16, device=DEV)
dbg = torch.zeros(256 * 16, dtype=torch.int64, device=DEV)
g0 = gates.clone()
for it in range(2):
    gates.copy_(g0)
    lib.las_lstm_rec_tc_set_debug(dbg.data_ptr() if it == 1 else None)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(lib.las_lstm_rec_bwd_tc(dout.data_ptr(), gates.data_ptr(), dgb.data_ptr(), cs.data_ptr(), w_t.data_ptr(), lens_dev.data_ptr(), None, B, T, H, ndir, ws.data_ptr(), nbytes, st), 'bwd')
    e1.record(); torch.cuda.synchronize()
    print(f'run {it}: {e0.elapsed_time(e1) * 1e3 / T:.2f} us/step')
lib.las_lstm_rec_tc_set_debug(None)
d = dbg.cpu().numpy().reshape(256, 16).astype(np.float64)
names = {0: 'P flag seen', 1: 'P tma issued', 2: 'M full', 3: 'M committed', 4: 'E operands issued', 5: 'E tfull', 6: 'E partial parked',
         7: 'cluster sync 1', 11: 'E dsmem reduced', 12: 'cluster sync 2', 8: 'E pointwise+bar', 9: 'E released', 10: 'E fp32 stores'}
lo, hi = 20, min(T - 1, 200)
prev_rel = d[lo - 1:hi - 1, 9]
print('cycles relative to the release of the previous step (mean over steps %d..%d):' % (lo, hi))
for k in [0, 1, 2, 3, 4, 5, 6, 7, 11, 8, 9, 10]:
    print(f'  {names[k]:20s} {np.mean(d[lo:hi, k] - prev_rel):9.0f}')
print('step period: %.0f cycles' % np.mean(d[lo:hi, 9] - prev_rel))
