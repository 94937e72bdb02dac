"""Data-parallel gradient exchange for the LAS train step: bucketed all-reduce overlapped with backward.

The reference is single-process (no DDP, no NCCL).  Training shards by batch -- utterances are independent
end to end (no batch statistics; locked-dropout masks are per sample) -- so the only exchange step is the gradient
all-reduce (SUM) plus, for exact loss normalisation with ragged targets, a scalar all-reduce of the non-pad token count
that the caller does.  One process per GPU; torch.distributed (NCCL over NVLink / NVSwitch on the GPU box, gloo in the CPU
tests) is the transport.

Design:
  * gradients live permanently in flat fp32 bucket buffers; every p.grad is a view into its bucket, so autograd
    accumulates in place, zero_grad is one memset per bucket, the all-reduce needs no packing copies, and the fused
    optimizer sees stable pointers;
  * buckets follow the order backward produces gradients: Speller -> pLSTM[n-1] -> ... -> pLSTM[0] -> base LSTM;
  * a post-accumulate-grad hook counts a bucket's parameters down and fires its async all-reduce the moment the last
    one lands, so communication of the Speller / upper pyramid overlaps the BPTT of the layers below (encoder weight
    gradients that the backward overlap accumulates on its second stream report through `p._las_grad_ready` instead:
    the all-reduce is then issued from that stream, i.e. ordered behind the accumulation);
  * parameters that never receive a gradient (spell.attention.final_map.*) are excluded up front -- the
    classic DDP "unused parameter" trap;
  * gradient accumulation (the reference trainer's `accu_grad`, src/train.py:163-165): run every micro-batch but the last inside
    `with reducer.no_sync():` -- the hooks then neither count nor launch, gradients just accumulate in the buckets -- and the last
    one outside it, which arms the countdown again and reduces the accumulated sum once.  A hook that fires on an already reduced
    bucket (a second backward without no_sync / zero_grad) raises instead of silently mixing reduced and local gradients.
"""
from __future__ import annotations

import contextlib
from typing import Callable, Dict, List, Optional, Sequence

import torch
import torch.distributed as dist


def default_bucket_key(name: str) -> str:
    """Bucket id of a parameter name, in backward order."""
    if name.startswith('listen.pyramid.plstms.'):
        return 'pyramid.' + name.split('.')[3]
    if name.startswith('listen.base.'):
        return 'base'
    return 'spell'


class BucketedGradReducer:
    def __init__(self, named_params: Sequence, process_group=None, bucket_key: Callable[[str], str] = default_bucket_key,
                 exclude: Sequence[str] = ('final_map',), world_size: Optional[int] = None):
        self.pg = process_group
        self.world_size = world_size if world_size is not None else (dist.get_world_size(process_group) if dist.is_initialized() else 1)
        groups: Dict[str, List] = {}
        order: List[str] = []
        self.excluded = []
        for name, p in named_params:
            if not p.requires_grad:
                continue
            if any(e in name for e in exclude):
                self.excluded.append(name)
                continue
            k = bucket_key(name)
            if k not in groups:
                groups[k] = []
                order.append(k)
            groups[k].append((name, p))
        # named_parameters() runs listener-first; backward produces gradients in the reverse order
        self.bucket_names = list(reversed(order))
        self.buckets: List[torch.Tensor] = []
        self.members: List[List[torch.nn.Parameter]] = []
        self._pending: List[int] = []
        self._handles: List[Optional[object]] = []
        self._bucket_of: Dict[int, int] = {}
        self._hooks = []
        self._sync = True
        for bi, k in enumerate(self.bucket_names):
            ps = [p for _, p in groups[k]]
            n = sum(p.numel() for p in ps)
            flat = torch.zeros(n, dtype=ps[0].dtype, device=ps[0].device)
            off = 0
            for p in ps:
                p.grad = flat[off:off + p.numel()].view_as(p)
                # gradients of this parameter may be accumulated in place outside autograd (functional.py, backward overlap); the code
                # that does so calls p._las_grad_ready(p) afterwards -- on the stream the accumulation ran on -- in place of the hook
                p._las_bucketed = True
                p._las_deferred = False
                p._las_grad_ready = self._make_hook(bi, from_autograd=False)
                off += p.numel()
                self._bucket_of[id(p)] = bi
                self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook(bi)))
            self.buckets.append(flat)
            self.members.append(ps)
            self._pending.append(len(ps))
            self._handles.append(None)

    def _make_hook(self, bi: int, from_autograd: bool = True):
        def hook(param):
            if from_autograd:
                # torch calls the post-accumulate-grad hook even when the incoming gradient is None (verified on torch 2.11): a layer
                # that took the backward-overlap route returned None for this parameter and will accumulate into p.grad LATER, on its
                # second stream, then report through p._las_grad_ready.  Counting this early call would launch the bucket's all-reduce
                # before the gradient exists (every rank would keep only its local gradient).
                if getattr(param, '_las_deferred', False):
                    return
            else:
                param._las_deferred = False
            if not self._sync:               # inside no_sync(): accumulate locally, reduce with the last micro-batch
                return
            if self._pending[bi] <= 0:
                raise RuntimeError('BucketedGradReducer: a gradient arrived for a bucket that was already reduced in this step; wrap '
                                   'all but the last backward of an accumulation cycle in `with reducer.no_sync():` (or call '
                                   'reducer.zero_grad() between steps)')
            self._pending[bi] -= 1
            if self._pending[bi] == 0:
                self._launch(bi)
        return hook

    @contextlib.contextmanager
    def no_sync(self):
        """Gradient accumulation: backward passes inside this context add into the buckets without communicating."""
        prev, self._sync = self._sync, False
        try:
            yield
        finally:
            self._sync = prev

    def _launch(self, bi: int):
        if self.world_size > 1 and self._handles[bi] is None:
            self._handles[bi] = dist.all_reduce(self.buckets[bi], op=dist.ReduceOp.SUM, group=self.pg, async_op=True)

    def zero_grad(self):
        """Replaces optimizer.zero_grad(): grads stay allocated (views of the buckets) and are zeroed in place."""
        for bi, flat in enumerate(self.buckets):
            flat.zero_()
            for p in self.members[bi]:
                if p.grad is None or p.grad.data_ptr() < flat.data_ptr() or p.grad.data_ptr() >= flat.data_ptr() + flat.numel() * flat.element_size():
                    raise RuntimeError('a parameter gradient was detached from its bucket (use reducer.zero_grad(), not '
                                       'optimizer.zero_grad(set_to_none=True))')
            self._pending[bi] = len(self.members[bi])
            self._handles[bi] = None

    def finish(self):
        """Call after backward: launches any bucket whose hooks did not all fire, then waits for every all-reduce.
        Gradients hold the SUM over ranks; fold 1/world_size into the optimizer's inv_scale."""
        for bi in range(len(self.buckets)):
            if self._handles[bi] is None:
                self._launch(bi)
        for bi, h in enumerate(self._handles):
            if h is not None:
                h.wait()

    @property
    def grad_bytes(self) -> int:
        return sum(b.numel() * b.element_size() for b in self.buckets)
