import sys, copy, numpy as np, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/attention-based-e2e-asr-dnn_b200')
from las_b200 import configs as gu
from las_b200.models import ListenAttendSpell
from las_b200.loss import masked_ce
dev='cuda:0'
cfg = gu.get_config('best')
Bl, T, L = 8, 160, 10
lens = [160, 152, 144, 131, 120, 97, 80, 64]
x, lx, y = gu.make_inputs(100, Bl, T, L, lens)
x, y, lx = torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev), torch.from_numpy(lx)
torch.manual_seed(5)
m0 = ListenAttendSpell(**copy.deepcopy(cfg)).to(dev).train()
def run(model, parts, amp):
    model.zero_grad(set_to_none=True)
    for sl in parts:
        xx, yy, ll = x[sl].contiguous(), y[sl].contiguous(), lx[sl]
        with torch.autocast('cuda', dtype=torch.bfloat16, enabled=amp):
            logits, _ = model(xx, ll, yy, 1.0, False)
        loss, _ = masked_ce(logits, yy, torch.full((xx.shape[0],), L, dtype=torch.int64))
        (loss * 256.0 * xx.shape[0] / Bl).backward()
    torch.cuda.synchronize()
    return {n: p.grad.detach().float().cpu().numpy().copy() for n, p in model.named_parameters() if p.grad is not None}
for amp in (False, True):
    full = run(m0, [slice(0, 8)], amp)
    for name, parts in (('4+4', [slice(0,4), slice(4,8)]), ('first4 twice vs', None), ('trimT', 'trim')):
        if parts is None or parts == 'trim':
            continue
        acc = run(m0, parts, amp)
        gmax = max(np.abs(v).max() for v in full.values())
        errs = sorted(((float(np.abs(acc[n]-full[n]).max()/max(np.abs(full[n]).max(), 1e-3*gmax)), n) for n in full), reverse=True)[:3]
        print('amp' if amp else 'fp32', name, errs)
    # same rows 4..7 alone, once with the tensor trimmed to T=120 and once padded to 160
    a = run(m0, [slice(4, 8)], amp)
    xt = x[4:8, :120].contiguous()
    m0.zero_grad(set_to_none=True)
    with torch.autocast('cuda', dtype=torch.bfloat16, enabled=amp):
        logits, _ = m0(xt, lx[4:8], y[4:8].contiguous(), 1.0, False)
    loss, _ = masked_ce(logits, y[4:8].contiguous(), torch.full((4,), L, dtype=torch.int64))
    (loss * 256.0 * 4 / Bl).backward(); torch.cuda.synchronize()
    b = {n: p.grad.detach().float().cpu().numpy().copy() for n, p in m0.named_parameters() if p.grad is not None}
    gmax = max(np.abs(v).max() for v in a.values())
    print('amp' if amp else 'fp32', 'rows4-7 padded-vs-trimmed', sorted(((float(np.abs(a[n]-b[n]).max()/max(np.abs(b[n]).max(),1e-3*gmax)), n) for n in a), reverse=True)[:2])
