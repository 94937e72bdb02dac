import os, sys, torch
sys.path.insert(0, '/root/repo/attention-based-e2e-asr-dnn_b200')
from las_b200 import functional as LF
dev = torch.device('cuda:0')
def timed(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
B, T, NG = 96, 1600, 4096
out = torch.empty(B, T, NG, device=dev)
ms = timed(lambda: out.zero_())
print(f'memset {out.numel()*4/1e9:.2f} GB: {ms:.3f} ms = {out.numel()*4/ms/1e6:.0f} GB/s')
src = torch.empty(B, T, NG // 2, device=dev)
dst = torch.empty_like(src)
ms = timed(lambda: dst.copy_(src))
print(f'copy {src.numel()*4/1e9:.2f} GB: {ms:.3f} ms = {2*src.numel()*4/ms/1e6:.0f} GB/s (read+write)')
for K in (64, 256, 1024, 2048):
    x = torch.randn(B * T, K, device=dev).to(torch.bfloat16)
    W = torch.randn(NG, K, device=dev).to(torch.bfloat16)
    b1 = torch.randn(NG, device=dev)
    ms = timed(lambda: LF.gemm_tc(x, W, out, T, NG, K, a_batches=B, a_s1=K, a_s2=T * K, b_s1=K, c_bs=T * NG, ldc=NG, bias1=b1, bias2=b1))
    print(f'gate GEMM M={B*T} N={NG} K={K}: {ms:.3f} ms = {2.0*B*T*NG*K/ms/1e9:.0f} TFLOP/s, output {out.numel()*4/ms/1e6:.0f} GB/s')
