"""Fused AdamW(amsgrad) with the GradScaler unscale and the global-norm clip folded in.

Replaces the sequence of reference src/train.py:165-183
    scaler.unscale_(optimizer); clip_grad_norm_(params, grad_norm); scaler.step(optimizer); scaler.update()
(~15 foreach launches, ~6 passes over the gradients) by two launches of las_adamw_amsgrad_fused.

Two ways to use it:
  * drop-in: `FusedAdamW(model.parameters(), lr=..., weight_decay=..., amsgrad=True)` is a torch.optim.Optimizer with
    the same state_dict layout as torch.optim.AdamW (step, exp_avg, exp_avg_sq, max_exp_avg_sq); `.step()` alone is the
    plain AdamW update (grads already unscaled / clipped by the caller, as train.py does).
  * fused: `.step_fused(inv_scale, max_norm)` does unscale + clip + update (or skips on non-finite grads, like
    GradScaler.step) and returns device-side (found_inf, grad_norm) without a host sync.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import List

import torch

from . import _lib
from ._lib import ADAM_CHUNK, LasAdamChunk, LasAdamTensor, check, stream_ptr


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, amsgrad=False):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=amsgrad)
        super().__init__(params, defaults)
        self._status = None
        self._pending = None          # (pinned host copy of [found_inf, norm], event, params that stepped) of the previous fused step

    def _init_state(self, p, amsgrad=True):
        st = self.state[p]
        if len(st) == 0:
            st['step'] = torch.tensor(0.0, dtype=torch.float32)      # same key set as torch.optim.AdamW
            st['exp_avg'] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.preserve_format)
        if amsgrad and 'max_exp_avg_sq' not in st:                  # like torch.optim.AdamW: only with amsgrad
            st['max_exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return st

    def _reconcile(self):
        """GradScaler.step skips the whole optimizer step on a non-finite gradient, so the step counters must not advance.  The
        kernel knows (status[0]); the host learns it one call later from a pinned copy whose event has long fired by then (a whole
        forward + backward was enqueued in between), and rolls the counters of the skipped step back -- exact bias correction
        without a host sync in the step that is being timed."""
        if self._pending is None:
            return
        host, ev, stepped = self._pending
        self._pending = None
        ev.synchronize()
        if float(host[0]) != 0.0:
            for p in stepped:
                self.state[p]['step'] -= 1

    def state_dict(self):
        self._reconcile()
        return super().state_dict()

    @torch.no_grad()
    def _run(self, inv_scale: float, max_norm: float):
        lib = _lib.load()
        self._reconcile()
        if max_norm > 0 and sum(1 for g in self.param_groups if any(p.grad is not None for p in g['params'])) > 1:
            # clip_grad_norm_(model.parameters()) is ONE norm over every parameter; the kernel computes it per call
            raise RuntimeError('FusedAdamW.step_fused clips by the global norm of ONE parameter group (the reference trainer builds one, '
                               'src/train.py:71-77); put the parameters in a single group, or clip yourself and call step()')
        stepped = []
        for group in self.param_groups:
            ps = [p for p in group['params'] if p.grad is not None]
            if not ps:
                continue
            dev = ps[0].device
            if not ps[0].is_cuda:
                raise RuntimeError('FusedAdamW runs on CUDA parameters only (no CPU fallback)')
            beta1, beta2 = group['betas']
            tab = (LasAdamTensor * len(ps))()
            chunks: List[tuple] = []
            for i, p in enumerate(ps):
                if p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError('FusedAdamW needs contiguous fp32 parameters')
                st = self._init_state(p, bool(group['amsgrad']))
                # like torch, the step counter advances before the update; when a non-finite gradient makes the kernel skip, it
                # is rolled back (see _reconcile / step_fused)
                st['step'] += 1
                stepped.append(p)
                step = float(st['step'])
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                if g.dtype != torch.float32:
                    g = g.float()
                p._las_g = g          # keep alive until the kernels ran
                t = tab[i]
                t.p, t.g, t.m, t.v = p.data_ptr(), g.data_ptr(), st['exp_avg'].data_ptr(), st['exp_avg_sq'].data_ptr()
                t.vmax = st['max_exp_avg_sq'].data_ptr() if group['amsgrad'] else 0
                t.numel = p.numel()
                t.step_size = group['lr'] / (1.0 - beta1 ** step)
                t.bias_c2_sqrt = math.sqrt(1.0 - beta2 ** step)
                for off in range(0, p.numel(), ADAM_CHUNK):
                    chunks.append((i, off))
            ck = (LasAdamChunk * len(chunks))()
            for j, (i, off) in enumerate(chunks):
                ck[j].tensor, ck[j].offset = i, off
            # pointer tables go up as one small pinned->device copy each (grads are re-allocated by autograd every step)
            tab_t = torch.frombuffer(bytearray(bytes(tab)), dtype=torch.uint8).to(dev, non_blocking=True)
            ck_t = torch.frombuffer(bytearray(bytes(ck)), dtype=torch.uint8).to(dev, non_blocking=True)
            scratch = torch.empty(len(chunks) + 8, dtype=torch.float32, device=dev)
            status = torch.empty(2, dtype=torch.float32, device=dev)
            check(lib.las_adamw_amsgrad_fused(tab_t.data_ptr(), len(ps), ck_t.data_ptr(), len(chunks), float(group['lr']),
                                              float(beta1), float(beta2), float(group['eps']), float(group['weight_decay']),
                                              float(inv_scale), float(max_norm), int(bool(group['amsgrad'])),
                                              scratch.data_ptr(), status.data_ptr(), stream_ptr()), 'adamw_amsgrad_fused')
            self._status = status
            self._keep = (tab_t, ck_t, scratch)
        if self._status is not None and stepped:
            host = torch.empty(2, dtype=torch.float32, pin_memory=True)
            host.copy_(self._status, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(stepped[0].device))
            self._pending = (host, ev, stepped)
        return self._status

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        self._run(1.0, 0.0)          # max_norm 0 => no clipping: grads were prepared by the caller
        return loss

    @torch.no_grad()
    def step_fused(self, inv_scale: float = 1.0, max_norm: float = 5.0, sync_skip: bool = False):
        """unscale + clip + AdamW in one go.  Returns the device tensor [found_inf, grad_norm].  A step skipped for a non-finite
        gradient does not advance the step counters (GradScaler.step semantics): with sync_skip=True the host reads found_inf now
        (one sync); without it the roll-back happens at the start of the next call / at state_dict() from a pinned copy (no sync)."""
        status = self._run(inv_scale, max_norm)
        if sync_skip:
            self._reconcile()
        return status
