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
            # like LSTMLayerFunction.backward: mark the parameters as deferred (autograd will still call their post-accumulate hook
            # with a None gradient), accumulate (here at once; in the product later, on the second stream), then report
            wp._las_deferred = bp._las_deferred = True
            ctx.deferred = (dy.t() @ x, dy.sum(0))
            _DEFERRED.append((wp, bp, ctx.deferred))
            return dx, None, None, None
        return dx, dy.t() @ x, dy.sum(0), None


_DEFERRED = []


def _flush_deferred():
    """Stands in for _BackwardOverlap.run_pending / _overlap_finish: the queued weight gradients are accumulated AFTER the autograd
    hooks of their parameters have already fired with None."""
    for wp, bp, (gw, gb) in _DEFERRED:
        wp.grad.add_(gw)
        bp.grad.add_(gb)
        for p in (wp, bp):
            p._las_grad_ready(p)
    _DEFERRED.clear()


class ToyManual(Toy):
    def forward(self, x):
        x = torch.tanh(self.listen.base(x))
        for l in self.listen.pyramid.plstms:
            x = torch.tanh(_ManualGradLinear.apply(x, l.weight, l.bias, (l.weight, l.bias)))
        return self.spell.cls(x)


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close(); return p


def _worker_accumulate(rank, world, port, q):
    """accu_grad = 2 (src/train.py:163-165): two micro-batches per optimizer step; the first inside no_sync()."""
    from las_b200.ddp import BucketedGradReducer
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(0)
    model = ToyManual()
    red = BucketedGradReducer(list(model.named_parameters()), world_size=world)
    g = torch.Generator().manual_seed(1)
    X = torch.randn(16, 5, generator=g)
    Y = torch.randn(16, 3, generator=g)
    xs, ys = X[rank * 8:(rank + 1) * 8], Y[rank * 8:(rank + 1) * 8]
    ok = True
    for it in range(2):
        red.zero_grad()
        with red.no_sync():
            ((model(xs[:4]) - ys[:4]) ** 2).sum().backward()
            _flush_deferred()
            ok = ok and all(h is None for h in red._handles)                 # nothing was communicated
        ((model(xs[4:]) - ys[4:]) ** 2).sum().backward()
        _flush_deferred()
        red.finish()
    # a further backward without no_sync / zero_grad must raise instead of mixing reduced and local gradients
    try:
        ((model(xs[:4]) - ys[:4]) ** 2).sum().backward()
        raised = False
    except RuntimeError as e:
        raised = 'already reduced' in str(e)
    _DEFERRED.clear()
    if rank == 0:
        q.put((bool(ok), bool(raised)))
    dist.destroy_process_group()


def _worker_accumulate_values(rank, world, port, q):
    from las_b200.ddp import BucketedGradReducer
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(0)
    model = ToyManual()
    red = BucketedGradReducer(list(model.named_parameters()), world_size=world)
    g = torch.Generator().manual_seed(1)
    X = torch.randn(16, 5, generator=g)
    Y = torch.randn(16, 3, generator=g)
    xs, ys = X[rank * 8:(rank + 1) * 8], Y[rank * 8:(rank + 1) * 8]
    for it in range(2):
        red.zero_grad()
        with red.no_sync():
            ((model(xs[:4]) - ys[:4]) ** 2).sum().backward()
            _flush_deferred()
        ((model(xs[4:]) - ys[4:]) ** 2).sum().backward()
        _flush_deferred()
        red.finish()
    grads = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
    if rank == 0:
        ref = Toy()
        ref.load_state_dict(model.state_dict())
        ((ref(X) - Y) ** 2).sum().backward()
        q.put(all(torch.allclose(grads[k], p.grad, atol=1e-5) for k, p in ref.named_parameters() if p.grad is not None))
    dist.destroy_process_group()


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
        if manual:
            # the None-gradient hook calls must NOT have launched the pyramid buckets: their gradients do not exist yet
            assert all(red._handles[red.bucket_names.index(k)] is None for k in ('pyramid.1', 'pyramid.0'))
            _flush_deferred()                # ... now they do, and the ready callbacks launch them, before finish()
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


def _run2(target):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=target, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    return q.get(timeout=5)


def test_gradient_accumulation_with_no_sync_matches_single_process():
    """Two micro-batches per step (the reference trainer's accu_grad): gradients equal one backward over the concatenated batch of
    both ranks; the first micro-batch communicates nothing."""
    assert _run2(_worker_accumulate_values) is True


def test_second_backward_without_no_sync_raises():
    assert _run2(_worker_accumulate) == (True, True)


class ToyMixedSpell(Toy):
    """The Speller's case (las_b200.functional.SpellerFunction.backward, deferred route): ONE bucket ('spell') holds parameters whose
    gradients arrive through autograd (key_map / value_map: ordinary Linear layers) and parameters whose gradients are accumulated later
    on the second stream and reported through p._las_grad_ready (the 13 parameters of the decoder loop; here: spell.cls)."""

    def __init__(self):
        super().__init__()
        self.spell.attention.key_map = torch.nn.Linear(6, 6)

    def forward(self, x):
        x = torch.tanh(self.listen.base(x))
        for l in self.listen.pyramid.plstms:
            x = torch.tanh(l(x))
        x = torch.tanh(self.spell.attention.key_map(x))
        c = self.spell.cls
        return _ManualGradLinear.apply(x, c.weight, c.bias, (c.weight, c.bias))


class ToyMixedRef(Toy):
    def __init__(self):
        super().__init__()
        self.spell.attention.key_map = torch.nn.Linear(6, 6)

    def forward(self, x):
        x = torch.tanh(self.listen.base(x))
        for l in self.listen.pyramid.plstms:
            x = torch.tanh(l(x))
        x = torch.tanh(self.spell.attention.key_map(x))
        return self.spell.cls(x)


def _worker_mixed_bucket(rank, world, port, q):
    from las_b200.ddp import BucketedGradReducer
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(0)
    model = ToyMixedSpell()
    red = BucketedGradReducer(list(model.named_parameters()), world_size=world)
    bi = red.bucket_names.index('spell')
    g = torch.Generator().manual_seed(1)
    X = torch.randn(8, 5, generator=g)
    Y = torch.randn(8, 3, generator=g)
    xs, ys = X[rank * 4:(rank + 1) * 4], Y[rank * 4:(rank + 1) * 4]
    ok = True
    for it in range(2):
        red.zero_grad()
        ((model(xs) - ys) ** 2).sum().backward()
        # key_map's gradients have arrived through autograd, the deferred ones have not: the bucket must NOT have been reduced yet
        # (every other bucket is complete and already in flight)
        ok = ok and red._handles[bi] is None and red._pending[bi] == 2
        ok = ok and all(red._handles[i] is not None for i in range(len(red.buckets)) if i != bi)
        _flush_deferred()
        ok = ok and red._handles[bi] is not None
        red.finish()
    grads = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
    if rank == 0:
        ref = ToyMixedRef()
        ref.load_state_dict(model.state_dict())
        ((ref(X) - Y) ** 2).sum().backward()
        same = all(torch.allclose(grads[k], p.grad, atol=1e-5) for k, p in ref.named_parameters() if p.grad is not None)
        q.put((bool(ok), bool(same)))
    dist.destroy_process_group()


def test_bucket_with_deferred_and_autograd_parameters():
    """The 'spell' bucket is reduced once, after BOTH its autograd-delivered gradients and the ones accumulated out of band exist."""
    assert _run2(_worker_mixed_bucket) == (True, True)
