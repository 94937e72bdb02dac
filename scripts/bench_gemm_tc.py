"""Per-shape timing of the tcgen05 gate GEMM (forward / dgrad / wgrad forms) at the best config's layer shapes.
CUDA events, L2 flushed between iterations, 3 warm-up + 10 timed."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'attention-based-e2e-asr-dnn_b200'))
import torch
from las_b200 import functional as LF
DEV = 'cuda:0'
B = int(os.environ.get('B', 96))
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=DEV)

def timeit(fn, n=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]

def rnd(*shape):
    return (torch.randn(*shape, device=DEV) * 0.1).to(torch.bfloat16)

H = 512; NG = 8 * H
for name, T, Din in [('L1 (T=800)', 800, 2048), ('L2 (T=400)', 400, 2048), ('L3 (T=200)', 200, 2048), ('L0 (T=1600,K=64 pad)', 1600, 64)]:
    x = rnd(B, T, Din); w = rnd(NG, Din); g = torch.empty(B, T, NG, device=DEV)
    dg = rnd(B * T, NG); dx = torch.empty(B, T, Din, device=DEV); dw = torch.empty(NG, Din, device=DEV)
    hs = rnd(B, T + 2, 2 * H); dwhh = torch.empty(4 * H, H, device=DEV)
    fl = 2.0 * B * T * NG * Din
    t = timeit(lambda: LF.gemm_tc(x, w, g, T, NG, Din, a_batches=B, a_s1=Din, a_s2=T * Din, b_s1=Din, c_bs=T * NG, ldc=NG))
    print(f'{name:22s} fwd   M={B*T:6d} N={NG} K={Din:4d}: {t*1e3:8.1f} us  {fl/t/1e9:7.1f} TFLOP/s')
    t = timeit(lambda: LF.gemm_tc(dg, w, dx, T, Din, NG, a_batches=B, a_s1=NG, a_s2=T * NG, b_s1=Din, b_mn=True, c_bs=T * Din, ldc=Din))
    print(f'{name:22s} dgrad M={B*T:6d} N={Din:4d} K={NG}: {t*1e3:8.1f} us  {fl/t/1e9:7.1f} TFLOP/s')
    t = timeit(lambda: LF.gemm_tc(dg, x, dw, NG, Din, T, k_batches=B, a_s1=NG, a_s2=T * NG, b_s1=Din, b_s2=T * Din, ldc=Din, a_mn=True, b_mn=True))
    print(f'{name:22s} wgrad M={NG} N={Din:4d} K={B*T:6d}: {t*1e3:8.1f} us  {fl/t/1e9:7.1f} TFLOP/s')
    fl2 = 2.0 * B * T * 4 * H * H
    t = timeit(lambda: LF.gemm_tc(dg, hs, dwhh, 4 * H, H, T, k_batches=B, a_s1=NG, a_s2=T * NG, b_s1=2 * H, b_s2=(T + 2) * 2 * H, ldc=H, a_mn=True, b_mn=True))
    print(f'{name:22s} dWhh  M={4*H} N={H:4d} K={B*T:6d}: {t*1e3:8.1f} us  {fl2/t/1e9:7.1f} TFLOP/s')
# reference point: cuBLAS bf16 on the same forward shape
x = rnd(B * 800, 2048); w = rnd(NG, 2048)
t = timeit(lambda: torch.matmul(x, w.t()))
print(f'cuBLAS bf16 (bf16 out) M={B*800} N={NG} K=2048: {t*1e3:8.1f} us  {2.0*B*800*NG*2048/t/1e9:7.1f} TFLOP/s')
