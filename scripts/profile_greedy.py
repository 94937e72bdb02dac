"""Greedy decode timing at BASELINE configs[3] (best config, B=256, T_enc=375, 600 steps): decoder loop only (Speller.forward in eval mode).
    python scripts/profile_greedy.py [B] [T_enc]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'attention-based-e2e-asr-dnn_b200'))
from las_b200 import configs
from las_b200.models import ListenAttendSpell
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
T = int(sys.argv[2]) if len(sys.argv) > 2 else 375
dev = torch.device('cuda:0')
torch.manual_seed(1)
model = ListenAttendSpell(**configs.get_config('best')).to(dev).eval()
enc_h = torch.randn(B, T, 1024, device=dev) * 0.3
enc_l = torch.full((B,), T, dtype=torch.int64)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for it in range(3):
    ev0.record()
    with torch.inference_mode(), torch.autocast('cuda', dtype=torch.bfloat16):
        logits, _ = model.spell(enc_h, enc_l)
    ev1.record()
    torch.cuda.synchronize()
    print(f'greedy decoder loop B={B} T_enc={T}: {ev0.elapsed_time(ev1):.2f} ms for 600 steps = {ev0.elapsed_time(ev1) * 1e3 / 600:.1f} us/step '
          f'(LAS_KV_F16={os.environ.get("LAS_KV_F16", "0")}, LAS_DP_KRES={os.environ.get("LAS_DP_KRES", "auto")})')
