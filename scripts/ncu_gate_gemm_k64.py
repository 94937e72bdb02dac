"""One launch shape for Nsight Compute: the base layer's gate projection (M = 96 x 1600, N = 4096, K = 64: all epilogue, 2.5 GB of fp32
gate pre-activations written).  ncu --set full --import-source on -k regex:gemm_bf16_tc -s 2 -c 1 python scripts/ncu_gate_gemm_k64.py"""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'attention-based-e2e-asr-dnn_b200'))
from las_b200 import functional as LF
dev = torch.device('cuda:0')
B, T, NG, K = 96, 1600, 4096, 64
out = torch.empty(B, T, NG, device=dev)
x = torch.randn(B * T, K, device=dev).to(torch.bfloat16)
W = torch.randn(NG, K, device=dev).to(torch.bfloat16)
b1 = torch.randn(NG, device=dev)
for _ in range(4):
    LF.gemm_tc(x, W, out, T, NG, K, a_batches=B, a_s1=K, a_s2=T * K, b_s1=K, c_bs=T * NG, ldc=NG, bias1=b1, bias2=b1)
torch.cuda.synchronize()
print('ok')
