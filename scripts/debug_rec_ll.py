"""Forward tensor-pipe recurrence vs the fp32 kernel: where (batch slice, direction, time) do they differ?"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'attention-based-e2e-asr-dnn_b200'))
import numpy as np, torch
from las_b200 import _lib, functional as LF
lib = _lib.load(); DEV = 'cuda:0'
H, B, T = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
rng = np.random.default_rng(H + B)
lens = [T] + [int(v) for v in rng.integers(1, T + 1, size=B - 1)]
ndir, F = 2, 2 * H
gates0 = torch.from_numpy(rng.standard_normal((B, T, ndir, 4 * H)).astype(np.float32)).to(DEV)
w_hh = torch.from_numpy((rng.uniform(-1, 1, size=(ndir, 4 * H, H)) / np.sqrt(H)).astype(np.float32)).to(DEV)
lens_dev = torch.tensor(lens, dtype=torch.int32, device=DEV)
res = []
for tc in (False, True):
    gates = gates0.clone()
    hs = torch.full((B, T + 2, F), 7.0, device=DEV); cs = torch.full((B, T + 2, F), 7.0, device=DEV); out = torch.full((B, T, F), 7.0, device=DEV)
    st = torch.cuda.current_stream().cuda_stream
    if tc:
        wb = LF.cast_bf16(w_hh, ndir * 4 * H, H, H, H)
        nbytes = lib.las_lstm_rec_tc_workspace_bytes(B, H, ndir); ws = torch.empty(nbytes, dtype=torch.uint8, device=DEV)
        _lib.check(lib.las_lstm_rec_fwd_tc(gates.data_ptr(), wb.data_ptr(), lens_dev.data_ptr(), None, out.data_ptr(), hs.data_ptr(), cs.data_ptr(), B, T, H, ndir, 1, ws.data_ptr(), nbytes, st), 'tc')
    else:
        nbytes = lib.las_lstm_rec_workspace_bytes(B, H, ndir); ws = torch.empty(nbytes, dtype=torch.uint8, device=DEV)
        _lib.check(lib.las_lstm_rec_fwd_f32(gates.data_ptr(), w_hh.data_ptr(), lens_dev.data_ptr(), None, out.data_ptr(), hs.data_ptr(), cs.data_ptr(), B, T, H, ndir, ws.data_ptr(), nbytes, st), 'f32')
    torch.cuda.synchronize()
    res.append(out)
d = (res[1] - res[0]).abs().cpu().numpy()
print('max err', d.max())
for sl in range((B + 31) // 32):
    for dr in range(2):
        e = d[sl * 32:(sl + 1) * 32, :, dr * H:(dr + 1) * H]
        print(f'slice {sl} dir {dr}: max {e.max():.4f} per-t', np.round(e.max(axis=(0, 2)), 3)[:12], 'per unit-block', np.round(e.reshape(e.shape[0], e.shape[1], -1, 32).max(axis=(0, 1, 3)), 3))
