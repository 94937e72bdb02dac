"""Per-phase timeline of the tensor-pipe recurrence forward (clock64 stamps of CTA (0,0,0))."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'attention-based-e2e-asr-dnn_b200'))
import numpy as np, torch
from las_b200 import _lib, functional as LF
lib = _lib.load()
DEV = 'cuda:0'
H, B, T = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
ndir, F = 2, 2 * H
rng = np.random.default_rng(0)
gates = torch.from_numpy(rng.standard_normal((B, T, ndir, 4 * H)).astype(np.float32)).to(DEV)
w_hh = torch.from_numpy((rng.uniform(-1, 1, size=(ndir, 4 * H, H)) / np.sqrt(H)).astype(np.float32)).to(DEV)
lens_dev = torch.full((B,), T, dtype=torch.int32, device=DEV)
hs = torch.zeros(B, T + 2, F, device=DEV); cs = torch.zeros(B, T + 2, F, device=DEV)
wb = LF.cast_bf16(w_hh, ndir * 4 * H, H, H, H)
nbytes = lib.las_lstm_rec_tc_workspace_bytes(B, H, ndir); ws = torch.zeros(nbytes, dtype=torch.uint8, device=DEV)
dbg = torch.zeros(256 * 16, dtype=torch.int64, device=DEV)
st = torch.cuda.current_stream().cuda_stream
for it in range(2):
    lib.las_lstm_rec_tc_set_debug(dbg.data_ptr() if it == 1 else None)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(lib.las_lstm_rec_fwd_tc(gates.data_ptr(), wb.data_ptr(), lens_dev.data_ptr(), None, None, hs.data_ptr(), cs.data_ptr(), B, T, H, ndir, int(os.environ.get("SAVE", "1")), ws.data_ptr(), nbytes, st), 'tc')
    e1.record(); torch.cuda.synchronize()
    print(f'run {it}: {e0.elapsed_time(e1) * 1e3 / T:.2f} us/step')
lib.las_lstm_rec_tc_set_debug(None)
d = dbg.cpu().numpy().reshape(256, 16).astype(np.float64)
names = ['P flag seen', 'P tma issued', 'M full', 'M committed', 'E start', 'E tfull', 'E tmem ld', 'E act+bar', 'E cell+bar', 'E released', 'E stores', 'P fence done']
if os.environ.get('LAS_REC_LL', '1') != '0':
    # LL exchange variant: slot 1 = loader: whole tile in smem (stamped during the step that consumes it), 2 = first k-block handed
    # to the MMA thread, 3 = MMAs committed, 9 = h_t published (tagged stores issued)
    names = ['L sentinels ok', 'L arrived', 'M tile ready', 'M committed', 'E start', 'E tfull', 'E tmem ld', 'E act+bar', '-', 'E published', 'E stores', 'L words valid']
lo, hi = 20, min(T - 1, 200)
base = d[lo:hi, 4]                      # epilogue start of step s
prev_rel = d[lo - 1:hi - 1, 9]          # release of step s-1
print('cycles relative to the release of the previous step (mean over steps %d..%d):' % (lo, hi))
for k, nm in enumerate(names):
    if nm == '-':
        continue
    print(f'  {nm:14s} {np.mean(d[lo:hi, k] - prev_rel):9.0f}')
print('step period (release to release): %.0f cycles' % np.mean(d[lo:hi, 9] - prev_rel))
