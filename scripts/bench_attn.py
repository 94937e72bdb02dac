"""Timing of the fused attention step kernels (fwd, bwd) at train (B=96,T=200) and greedy (B=256,T=375) shapes.

`python scripts/bench_attn.py`            sweeps LAS_ATTN_SPLIT (CTAs per batch row; 0 = two-phase one-CTA-per-row kernel)
`python scripts/bench_attn.py one`        times the current environment's setting only
Warm = K/V re-read by back-to-back launches (what the decoder loop does: K/V of B=96 fit in L2); cold = L2 flushed first.
"""
import sys, os, subprocess
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'attention-based-e2e-asr-dnn_b200'))


def one():
    import ctypes as C
    import torch
    from las_b200 import _lib
    from las_b200._lib import LasAttnStep
    lib = _lib.load()
    DEV = 'cuda:0'
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=DEV)
    st = torch.cuda.current_stream().cuda_stream
    for B, T, P in [(96, 200, 256), (256, 375, 256)]:
        q = torch.randn(B, P, device=DEV); K = torch.randn(B, T, P, device=DEV); V = torch.randn(B, T, P, device=DEV)
        lens = torch.full((B,), T, dtype=torch.int32, device=DEV)
        ctx = torch.empty(B, P, device=DEV); w = torch.empty(B, 1, T, device=DEV)
        dctx = torch.randn(B, P, device=DEV); dq = torch.empty(B, P, device=DEV); de = torch.empty(B, 1, T, device=DEV)
        d = LasAttnStep()
        d.q, d.ld_q = q.data_ptr(), P
        d.K, d.V, d.lens = K.data_ptr(), V.data_ptr(), lens.data_ptr()
        d.w, d.ld_w = w.data_ptr(), T
        d.ctx, d.ld_ctx = ctx.data_ptr(), P
        d.dctx, d.ld_dctx = dctx.data_ptr(), P
        d.dq, d.ld_dq, d.dq_accumulate = dq.data_ptr(), P, 0
        d.de = de.data_ptr()
        d.B, d.T, d.P, d.heads = B, T, P, 1
        d.scale = float(P ** 0.5)
        byts = 2.0 * B * T * P * 4
        for name, fn in (('fwd', lib.las_attn_step_fwd_f32), ('bwd', lib.las_attn_step_bwd_f32)):
            # warm: 50 launches captured in a CUDA graph and replayed (how the decoder loop issues them: no host gap)
            g = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                _lib.check(fn(C.byref(d), side.cuda_stream), name)
                torch.cuda.synchronize()
                with torch.cuda.graph(g, stream=side):
                    for _ in range(50):
                        _lib.check(fn(C.byref(d), side.cuda_stream), name)
            torch.cuda.synchronize()
            for mode in ('warm', 'cold'):
                ts = []
                for i in range(13):
                    NREP = 1 if mode == 'cold' else 50
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    if mode == 'cold':
                        flush.zero_()
                        e0.record()
                        _lib.check(fn(C.byref(d), st), name)
                        e1.record()
                    else:
                        e0.record()
                        g.replay()
                        e1.record()
                    torch.cuda.synchronize()
                    if i >= 3:
                        ts.append(e0.elapsed_time(e1) / NREP)
                ts.sort(); t = ts[len(ts) // 2]
                print(f'  attn {name} B={B} T={T} {mode}: {t*1e3:7.1f} us  {byts/t/1e6:8.1f} GB/s (algorithmic {byts/1e6:.1f} MB)', flush=True)


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'one':
        one()
    else:
        for s in (0, 1, 2, 3, 4, 6, 8):
            print(f'LAS_ATTN_SPLIT={s}', flush=True)
            env = dict(os.environ, LAS_ATTN_SPLIT=str(s))
            r = subprocess.run([sys.executable, __file__, 'one'], env=env, capture_output=True, text=True)
            print(r.stdout, end='')
            if r.returncode:
                print('  FAILED:', r.stderr[-400:])
