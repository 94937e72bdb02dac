"""Two-GPU NCCL test of the REAL data-parallel path: best-config dims, bf16 mode, backward overlap on
(encoder weight gradients accumulated on the second stream and reported through the reducer's ready callback, bucket all-reduces
issued from that stream), gradient accumulation with no_sync -- against a single-process run on the concatenated batch through plain
autograd.  Skipped with fewer than two GPUs (run with `gpurun --gpus 2`)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'attention-based-e2e-asr-dnn_b200')


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q, accumulate):
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    import copy
    import torch.distributed as dist
    from las_b200 import configs as gu
    from las_b200.ddp import BucketedGradReducer
    from las_b200.loss import masked_ce
    from las_b200.models import ListenAttendSpell
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    cfg = gu.get_config('best')
    Bl, T, L = 8, 160, 10                          # per-rank batch; ragged input lengths, full target lengths
    lens = [160, 152, 144, 131, 120, 97, 80, 64]
    xs, lxs, ys = [], [], []
    for r in range(world):
        x, lx, y = gu.make_inputs(100 + r, Bl, T, L, lens)
        xs.append(x); lxs.append(lx); ys.append(y)
    torch.manual_seed(5)
    model = ListenAttendSpell(**copy.deepcopy(cfg)).to(dev).train()
    ref_model = copy.deepcopy(model) if rank == 0 else None
    red = BucketedGradReducer(list(model.named_parameters()), world_size=world)
    ly = torch.full((Bl,), L, dtype=torch.int64)
    x, y, lx = torch.from_numpy(xs[rank]).to(dev), torch.from_numpy(ys[rank]).to(dev), torch.from_numpy(lxs[rank])

    def fwd_bwd(m, xx, lxx, yy, lyy):
        with torch.autocast('cuda', dtype=torch.bfloat16):
            logits, _ = m(xx, lxx, yy, 1.0, False)
        loss, _ = masked_ce(logits, yy, lyy)
        (loss * 256.0).backward()

    for it in range(2):                            # second iteration: re-armed hooks, pooled side stream
        red.zero_grad()
        if accumulate:                             # two micro-batches of 4 rows, the first without communication
            with red.no_sync():
                fwd_bwd(model, x[:4].contiguous(), lx[:4], y[:4].contiguous(), ly[:4])
            fwd_bwd(model, x[4:].contiguous(), lx[4:], y[4:].contiguous(), ly[4:])
        else:
            fwd_bwd(model, x, lx, y, ly)
        red.finish()
    torch.cuda.synchronize()
    got = {n: p.grad.detach().float().cpu().numpy() for n, p in model.named_parameters() if p.grad is not None}
    if rank == 0:
        # single process, concatenated batch, plain autograd (no reducer -> no overlap)
        xc = torch.from_numpy(np.concatenate(xs, 0)).to(dev)
        yc = torch.from_numpy(np.concatenate(ys, 0)).to(dev)
        lxc = torch.from_numpy(np.concatenate(lxs, 0))
        if accumulate:
            # the accumulated sum of two half-batch means == 2 x the mean over the whole rank batch, per rank
            scale = 0.25
        else:
            scale = 0.5
        fwd_bwd(ref_model, xc, lxc, yc, torch.full((world * Bl,), L, dtype=torch.int64))
        torch.cuda.synchronize()
        worst = ('', 0.0)
        errs = []
        gmax = max(float(p.grad.abs().max()) for p in ref_model.parameters() if p.grad is not None)
        for n, p in ref_model.named_parameters():
            if p.grad is None:
                continue
            ref = p.grad.detach().float().cpu().numpy()
            e = float(np.abs(got[n] * scale - ref).max() / max(float(np.abs(ref).max()), 1e-3 * gmax))
            errs.append((e, n))
            if e > worst[1]:
                worst = (n, e)
        print('largest errors:', sorted(errs, reverse=True)[:6], flush=True)
        q.put(worst)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs (gpurun --gpus 2)')
@pytest.mark.parametrize('accumulate', [False, True], ids=['one_backward', 'accu_grad_2'])
def test_allreduced_gradients_equal_the_concatenated_batch(accumulate):
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, accumulate)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=600)
        assert p.exitcode == 0
    name, err = q.get(timeout=10)
    print(f'2-GPU all-reduced gradients vs concatenated batch: worst {name} rel err {err:.2e}')
    # same bf16 operand roundings per utterance in both runs (rows are independent); only fp32 summation order over the batch differs
    assert err < 2e-3, (name, err)
