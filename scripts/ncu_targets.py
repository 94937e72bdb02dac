"""Launches exactly the kernels we want `ncu --set full` evidence for, at the bench shapes:
  1. gate GEMM forward, layer 1 (M=76800, N=4096, K=2048) + dgrad + wgrad
  2. fused attention step forward at the train (B=96,T=200) and greedy (B=256,T=375) shapes
  3. tensor-pipe recurrence forward + BPTT (B=96, H=512, T=200)
"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'attention-based-e2e-asr-dnn_b200'))
import numpy as np, torch
from las_b200 import _lib, functional as LF
lib = _lib.load(); DEV = 'cuda:0'; st = torch.cuda.current_stream().cuda_stream
REP = 1 if os.environ.get('NCU_CAPTURE') else 2          # under ncu every launch is replayed ~40 times: one pass is enough
def rnd(*s): return (torch.randn(*s, device=DEV) * 0.1).to(torch.bfloat16)
B, T, Din, H = 96, 800, 2048, 512; NG = 8 * H
x = rnd(B, T, Din); w = rnd(NG, Din); g = torch.empty(B, T, NG, device=DEV); dg = rnd(B * T, NG); dx = torch.empty(B, T, Din, device=DEV); dw = torch.empty(NG, Din, device=DEV)
for _ in range(REP):
    LF.gemm_tc(x, w, g, T, NG, Din, a_batches=B, a_s1=Din, a_s2=T * Din, b_s1=Din, c_bs=T * NG, ldc=NG)
    LF.gemm_tc(dg, w, dx, T, Din, NG, a_batches=B, a_s1=NG, a_s2=T * NG, b_s1=Din, b_mn=True, c_bs=T * Din, ldc=Din)
    LF.gemm_tc(dg, x, dw, NG, Din, T, k_batches=B, a_s1=NG, a_s2=T * NG, b_s1=Din, b_s2=T * Din, ldc=Din, a_mn=True, b_mn=True)
for (Ba, Ta) in [(96, 200), (256, 375)]:
    q = torch.randn(Ba, 256, device=DEV); K = torch.randn(Ba, Ta, 256, device=DEV); V = torch.randn(Ba, Ta, 256, device=DEV)
    lens = torch.full((Ba,), Ta, dtype=torch.int32, device=DEV)
    for _ in range(REP):
        LF.AttnStepFunction.apply(q, K, V, lens, 1)
Tr = 200; ndir = 2; F = 2 * H
rng = np.random.default_rng(0)
gates = torch.randn(B, Tr, ndir, 4 * H, device=DEV); w_hh = (torch.rand(ndir, 4 * H, H, device=DEV) * 2 - 1) / H ** 0.5
lens_dev = torch.full((B,), Tr, dtype=torch.int32, device=DEV)
hs = torch.zeros(B, Tr + 2, F, device=DEV); cs = torch.zeros(B, Tr + 2, F, device=DEV); dout = torch.randn(B, Tr, F, device=DEV)
wb = LF.cast_bf16(w_hh, ndir * 4 * H, H, H, H)
nbytes = lib.las_lstm_rec_tc_workspace_bytes(B, H, ndir); ws = torch.zeros(nbytes, dtype=torch.uint8, device=DEV)
w_t = torch.empty(ndir, H, 4 * H, dtype=torch.bfloat16, device=DEV)
_lib.check(lib.las_transpose_cast_bf16(w_hh.data_ptr(), w_t.data_ptr(), ndir, 4 * H, H, st), 'tr')
dgb = torch.empty(B * Tr, ndir * 4 * H, dtype=torch.bfloat16, device=DEV)
_lib.check(lib.las_lstm_rec_fwd_tc(gates.data_ptr(), wb.data_ptr(), lens_dev.data_ptr(), None, None, hs.data_ptr(), cs.data_ptr(), B, Tr, H, ndir, 1, ws.data_ptr(), nbytes, st), 'fwd')
_lib.check(lib.las_lstm_rec_bwd_tc(dout.data_ptr(), gates.data_ptr(), dgb.data_ptr(), cs.data_ptr(), w_t.data_ptr(), lens_dev.data_ptr(), None, B, Tr, H, ndir, ws.data_ptr(), nbytes, st), 'bwd')
# 4. the persistent decoder-step kernel (Speller forward loop as one cooperative launch): best dims, B=96, T_enc=200, 60 steps
from las_b200 import configs
from las_b200.models import ListenAttendSpell
torch.manual_seed(1)
m = ListenAttendSpell(**configs.get_config('best', dec_lstm_dropout=0.3)).to(DEV).train()
enc_h = torch.randn(B, 200, 1024, device=DEV) * 0.3
yy = torch.randint(1, 29, (B, 60), device=DEV)
for _ in range(REP):
    with torch.autocast('cuda', dtype=torch.bfloat16):
        m.spell(enc_h, torch.full((B,), 200, dtype=torch.int64), yy, 1.0, False)
# 5. six decoder steps forward + backward: the fused backward step (attention backward + dq.Wq + cell-1 backward, attn_step_split_kernel<1,0,1>)
enc_g = enc_h.clone().requires_grad_(True)
for _ in range(REP):
    with torch.autocast('cuda', dtype=torch.bfloat16):
        lg, _ = m.spell(enc_g, torch.full((B,), 200, dtype=torch.int64), yy[:, :6], 1.0, False)
    lg.float().square().mean().backward()
torch.cuda.synchronize(); print('ok')
