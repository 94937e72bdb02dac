"""World-size-2 gloo test (CPU) of the data-parallel gradient exchange (las_b200/ddp.py): bucket layout in backward
order, hook-driven async all-reduce, exclusion of never-used parameters, zero_grad semantics, and equality with a
single-process run on the concatenated batch."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'attention-based-e2e-asr-dnn_b200'))


class Toy(torch.nn.Module):
    """Same top-level naming scheme as ListenAttendSpell so default_bucket_key applies."""

    def __init__(self):
        super().__init__()
        self.listen = torch.nn.Module()
        self.listen.base = torch.nn.Linear(5, 7)
        self.listen.pyramid = torch.nn.Module()
        self.listen.pyramid.plstms = torch.nn.ModuleList([torch.nn.Linear(7, 7), torch.nn.Linear(7, 6)])
        self.spell = torch.nn.Module()
        self.spell.cls = torch.nn.Linear(6, 3)
        self.spell.attention = torch.nn.Module()
        self.spell.attention.final_map = torch.nn.Linear(3, 3)       # never used -> never gets a grad

    def forward(self, x):
        x = torch.tanh(self.listen.base(x))
        for l in self.listen.pyramid.plstms:
            x = torch.tanh(l(x))
        return self.spell.cls(x)


class _ManualGradLinear(torch.autograd.Function):
    """Stand-in for the encoder layers' backward overlap (las_b200/functional.py): parameters tagged by the reducer get their
    gradients accumulated straight into p.grad, report through p._las_grad_ready, and autograd sees None for them."""

    @staticmethod
    def forward(ctx, x, w, b, refs):
        ctx.save_for_backward(x, w)
        ctx.refs = refs
        return x @ w.t() + b

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        wp, bp = ctx.refs
        dx = dy @ w
        if getattr(wp, '_las_bucketed', False) and wp.grad is not None:
            wp.grad.add_(dy.t() @ x)
            bp.grad.add_(dy.sum(0))
            for p in (wp, bp):
                p._las_grad_ready(p)
            return dx, None, None, None
        return dx, dy.t() @ x, dy.sum(0), None


class ToyManual(Toy):
    def forward(self, x):
        x = torch.tanh(self.listen.base(x))
        for l in self.listen.pyramid.plstms:
            x = torch.tanh(_ManualGradLinear.apply(x, l.weight, l.bias, (l.weight, l.bias)))
        return self.spell.cls(x)


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q, manual=False):
    from las_b200.ddp import BucketedGradReducer
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(0)
    model = ToyManual() if manual else Toy()
    red = BucketedGradReducer(list(model.named_parameters()), world_size=world)
    assert red.bucket_names == ['spell', 'pyramid.1', 'pyramid.0', 'base'], red.bucket_names      # backward order
    assert red.excluded == ['spell.attention.final_map.weight', 'spell.attention.final_map.bias']
    g = torch.Generator().manual_seed(1)
    X = torch.randn(8, 5, generator=g)
    Y = torch.randn(8, 3, generator=g)
    xs, ys = X[rank * 4:(rank + 1) * 4], Y[rank * 4:(rank + 1) * 4]
    for it in range(2):                      # second iteration checks zero_grad / re-arming
        red.zero_grad()
        loss = ((model(xs) - ys) ** 2).sum()
        loss.backward()
        if manual:                           # the pyramid buckets were launched by the ready callbacks, before finish()
            assert all(red._handles[red.bucket_names.index(k)] is not None for k in ('pyramid.1', 'pyramid.0'))
        red.finish()
    grads = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
    # every p.grad is still a view into its bucket
    for b, ps in zip(red.buckets, red.members):
        for p in ps:
            assert b.data_ptr() <= p.grad.data_ptr() < b.data_ptr() + b.numel() * 4
    if rank == 0:
        ref = Toy()
        ref.load_state_dict(model.state_dict())
        ((ref(X) - Y) ** 2).sum().backward()
        ok = all(torch.allclose(grads[k], p.grad, atol=1e-5) for k, p in ref.named_parameters() if p.grad is not None)
        q.put(bool(ok) and ('spell.attention.final_map.weight' not in grads))
    dist.destroy_process_group()


@pytest.mark.parametrize('manual', [False, True], ids=['autograd', 'manual_accumulate'])
def test_bucketed_allreduce_matches_single_process(manual):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, manual)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
