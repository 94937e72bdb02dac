"""autograd.Function wrappers around the C ABI (include/las_b200.h).

PyTorch is plumbing here: it owns device memory (inputs, outputs, saved activations, workspaces), streams and the
autograd graph.  All arithmetic of the hot path runs in liblas_b200.so; there is no PyTorch/CPU fallback -- a CPU
tensor or a missing library raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import LasAttnStep, LasGemmF32, LasGemmTc, LasSpeller, LasSpellerGrads, check, ptr, stream_ptr
from .precision import use_tensor_cores


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError('las_b200 runs on CUDA tensors only (sm_100a kernels, no CPU fallback); got a '
                               f'{t.device} tensor')


def _f32c(t: torch.Tensor) -> torch.Tensor:
    """fp32 + contiguous view/copy (parameters already are)."""
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


def gemm_raw(A, B, Cout, M, N, K, *, am=(0, 0, 0), ak=(0, 1, 0), bk=(0, 1, 0), bn=1, cm=(0, 0, 0), bias1=None, bias2=None,
             alpha=1.0, beta=0.0, batch=1, bsA=0, bsB=0, bsC=0, a_off=0, b_off=0, c_off=0, gate=False):
    """Thin wrapper over las_gemm_f32.  am/ak/bk/cm = (s_outer, s_inner, inner); *_off are element offsets."""
    d = LasGemmF32()
    d.A = A.data_ptr() + 4 * a_off
    d.B = B.data_ptr() + 4 * b_off
    d.C = Cout.data_ptr() + 4 * c_off
    d.bias1 = ptr(bias1)
    d.bias2 = ptr(bias2)
    d.M, d.N, d.K, d.batch = int(M), int(N), int(K), int(batch)
    d.a_m_so, d.a_m_si, d.a_m_inner = int(am[0]), int(am[1]), int(am[2])
    d.a_k_so, d.a_k_si, d.a_k_inner = int(ak[0]), int(ak[1]), int(ak[2])
    d.b_k_so, d.b_k_si, d.b_k_inner = int(bk[0]), int(bk[1]), int(bk[2])
    d.b_n_s = int(bn)
    d.c_m_so, d.c_m_si, d.c_m_inner = int(cm[0]), int(cm[1]), int(cm[2])
    d.bsA, d.bsB, d.bsC = int(bsA), int(bsB), int(bsC)
    d.alpha, d.beta = float(alpha), float(beta)
    d.prof_tag = 1 if gate else 0
    check(_lib.load().las_gemm_f32(C.byref(d), stream_ptr()), 'gemm_f32')


def gemm_tc(A, B, Cout, M, N, K, *, a_batches=1, k_batches=1, a_s1, a_s2=0, b_s1, b_s2=0, c_bs=0, ldc, a_mn=False, b_mn=False,
            bias1=None, bias2=None, accumulate=False, lens=None, a_off=0, b_off=0, c_off=0, gate=True, flops=0.0, splitk=0,
            max_ctas=0, side=False, k_chunk=0, k_chunk_stride=0):
    """las_gemm_bf16_tc wrapper; A/B are bf16 (or fp16: the format is taken from the tensor's dtype) tensors, Cout fp32; *_off are
    element offsets.  k_chunk > 0: the reduction runs over K / k_chunk chunks whose starts are k_chunk_stride elements apart in both
    operands (las_b200.h)."""
    d = LasGemmTc()
    d.a_f16, d.b_f16 = int(A.dtype == torch.float16), int(B.dtype == torch.float16)
    d.A = A.data_ptr() + 2 * a_off
    d.B = B.data_ptr() + 2 * b_off
    d.C = Cout.data_ptr() + 4 * c_off
    d.bias1, d.bias2 = ptr(bias1), ptr(bias2)
    d.M, d.N, d.K = int(M), int(N), int(K)
    d.a_batches, d.k_batches = int(a_batches), int(k_batches)
    d.a_s1, d.a_s2, d.b_s1, d.b_s2 = int(a_s1), int(a_s2), int(b_s1), int(b_s2)
    d.c_bs, d.ldc = int(c_bs), int(ldc)
    d.a_mn_major, d.b_mn_major, d.accumulate = int(a_mn), int(b_mn), int(accumulate)
    d.lens = ptr(lens)
    d.prof_tag = (2 if side else 1) if gate else 0          # side: beside a recurrence kernel, timed apart (las_b200.h)
    d.prof_flops = float(flops)
    d.max_ctas = int(max_ctas)
    d.k_chunk, d.k_chunk_stride = int(k_chunk), int(k_chunk_stride)
    ws = None
    if a_mn and splitk == 0:
        # weight-gradient form: few output tiles, long reduction -> split K so every SM has a tile
        tiles = ((int(M) + 127) // 128) * ((int(N) + 255) // 256)
        if tiles < 74 and int(N) % 4 == 0:
            splitk = max(1, min(16, 148 // tiles, int(k_batches) * ((int(K) + 63) // 64)))
    if a_mn and splitk > 1:
        ws = torch.empty(splitk * int(M) * ((int(N) + 3) // 4 * 4), dtype=torch.float32, device=Cout.device)
        d.splitk, d.workspace = int(splitk), ws.data_ptr()
    check(_lib.load().las_gemm_bf16_tc(C.byref(d), stream_ptr()), 'gemm_bf16_tc')


def cast_bf16(src: torch.Tensor, rows: int, cols: int, cols_pad: int, ld_src: int, *, inner: int = 0, bs: int = 0,
              dst: Optional[torch.Tensor] = None, dst_off: int = 0) -> torch.Tensor:
    """fp32 rows x cols -> bf16 rows x cols_pad (zero padded columns).  Source row r lives at
    (r // inner) * bs + (r % inner) * ld_src (inner == 0: r * ld_src).  Writes into `dst` (+dst_off elements) or a new tensor."""
    if dst is None:
        dst = torch.empty(rows, cols_pad, dtype=torch.bfloat16, device=src.device)
    check(_lib.load().las_cast_f32_to_bf16(src.data_ptr(), int(ld_src), int(inner), int(bs), dst.data_ptr() + 2 * dst_off, int(cols_pad),
                                           int(rows), int(cols), int(cols_pad), stream_ptr()), 'cast_f32_to_bf16')
    return dst


def colsum(X: torch.Tensor, ld: int, M: int, N: int, out: torch.Tensor, x_off: int = 0, accumulate: bool = False):
    lib = _lib.load()
    scratch = torch.empty(lib.las_colsum_scratch_floats(int(N)), dtype=torch.float32, device=X.device)
    check(lib.las_colsum_f32(X.data_ptr() + 4 * x_off, int(ld), int(M), int(N), out.data_ptr(), int(accumulate),
                             scratch.data_ptr(), stream_ptr()), 'colsum')


def _rows3d(x: torch.Tensor):
    """(B, T, F) tensor with contiguous features -> (batch stride, time stride)."""
    assert x.dim() == 3 and x.stride(2) == 1
    return x.stride(0), x.stride(1)


def _attach_bf16_shadow(y: torch.Tensor, y16: Optional[torch.Tensor]) -> None:
    """Leaves the kernel-written bf16 copy on the fp32 tensor for the next lstm_layer / linear (see _bf16_shadow).  Inference
    tensors (the reference evaluates under torch.inference_mode(), src/train.py:207,604) have no version counter, so a later
    in-place edit could not be detected: no copy is attached there and the consumer casts as before."""
    if y16 is not None and not y.is_inference():
        y._las_bf16 = (y16, y._version)


def _bf16_shadow(x: torch.Tensor) -> Optional[torch.Tensor]:
    """bf16 copy of `x` left on it by the recurrence kernel that produced it (lstm_layer), or None.  Only valid while `x` has not
    been written to since (version counter), so an in-place edit by the caller silently falls back to the cast pass."""
    sh = getattr(x, '_las_bf16', None)
    if sh is None or x.is_inference() or not use_tensor_cores():
        return None
    t, ver = sh
    if ver != x._version or tuple(t.shape) != tuple(x.shape) or t.device != x.device or not t.is_contiguous():
        return None
    return t


# ----------------------------------------------------------------------------------------------------------------------
# nn.Linear on (B, T, F) / (M, F) inputs -- key_map / value_map / query_map (reference src/models.py:143-149,166)
# ----------------------------------------------------------------------------------------------------------------------
class LinearFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, bias2=None, x16=None):
        _require_cuda(x, weight, bias, bias2)
        x = x if x.dtype == torch.float32 else x.float()
        if x.stride(-1) != 1:
            x = x.contiguous()
        weight = _f32c(weight)
        N, K = weight.shape
        if x.dim() == 3:
            Bn, T, _ = x.shape
            sb, st = _rows3d(x)
            am = (sb, st, T)
            M = Bn * T
            out = torch.empty(Bn, T, N, dtype=torch.float32, device=x.device)
        else:
            x2 = x.reshape(-1, K)
            if not x2.is_contiguous():
                x2 = x2.contiguous()
            x = x2
            M = x.shape[0]
            am = (0, K, 0)
            out = torch.empty(*x.shape[:-1], N, dtype=torch.float32, device=x.device)
        b1 = _f32c(bias) if bias is not None else None
        b2 = _f32c(bias2) if bias2 is not None else None
        # the arithmetic mode must not depend on the batch: a size threshold here (M >= 64) made key_map / value_map fall back to fp32 for a
        # short micro-batch and rounded differently from the same rows inside a larger batch (gradient accumulation test, 2.6e-3)
        tc = use_tensor_cores() and K % 8 == 0 and N % 4 == 0
        if tc:
            # AMP mode: bf16 operands on the tensor pipe, fp32 accumulate / output
            xb = x16.view(M, K) if x16 is not None else cast_bf16(x, M, K, K, am[1], inner=am[2], bs=am[0])   # (M, K) bf16 compact
            wb = cast_bf16(weight, N, K, K, K)
            gemm_tc(xb, wb, out, M, N, K, a_s1=K, b_s1=K, ldc=N, bias1=b1, bias2=b2, gate=False)
            ctx.save_for_backward(xb, wb)
        else:
            gemm_raw(x, weight, out, M, N, K, am=am, ak=(0, 1, 0), bk=(0, 1, 0), bn=K, cm=(0, N, 0), bias1=b1, bias2=b2)
            ctx.save_for_backward(x, weight)
        ctx.am, ctx.M, ctx.has_bias, ctx.has_bias2, ctx.tc, ctx.xshape = am, M, bias is not None, bias2 is not None, tc, tuple(x.shape)
        return out

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        N, K = weight.shape
        M, am = ctx.M, ctx.am
        dy = _f32c(dy)
        dx = dw = db = None
        if ctx.tc:
            dyb = cast_bf16(dy, M, N, N, N)
            if ctx.needs_input_grad[0]:
                dx = torch.empty(ctx.xshape, dtype=torch.float32, device=dy.device)
                gemm_tc(dyb, weight, dx, M, K, N, a_s1=N, b_s1=K, b_mn=True, ldc=K, gate=False)
            if ctx.needs_input_grad[1]:
                dw = torch.empty(N, K, dtype=torch.float32, device=dy.device)
                gemm_tc(dyb, x, dw, N, K, M, a_s1=N, b_s1=K, ldc=K, a_mn=True, b_mn=True, gate=False)
        else:
            if ctx.needs_input_grad[0]:
                dx = torch.empty(x.shape, dtype=torch.float32, device=x.device)
                gemm_raw(dy, weight, dx, M, K, N, am=(0, N, 0), ak=(0, 1, 0), bk=(0, K, 0), bn=1, cm=(0, K, 0))
            if ctx.needs_input_grad[1]:
                dw = torch.empty_like(weight)
                # dW[n][k] = sum_m dy[m][n] x[m][k]
                gemm_raw(dy, x, dw, N, K, M, am=(0, 1, 0), ak=(0, N, 0), bk=am, bn=1, cm=(0, K, 0))
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = torch.empty(N, dtype=torch.float32, device=dy.device)
            colsum(dy, N, M, N, db)
        db2 = None
        if ctx.has_bias2 and ctx.needs_input_grad[3]:
            if db is not None:
                db2 = db.clone()
            else:
                db2 = torch.empty(N, dtype=torch.float32, device=dy.device)
                colsum(dy, N, M, N, db2)
        return dx, dw, db, db2, None


def linear(x, weight, bias=None, bias2=None):
    return LinearFunction.apply(x, weight, bias, bias2, _bf16_shadow(x) if x.dim() == 3 else None)


# ----------------------------------------------------------------------------------------------------------------------
# Backward overlap: weight-gradient GEMMs of layer l+1 beside the BPTT kernel of layer l.
#
# The BPTT kernel is latency bound and occupies rs * slices * ndir SMs (96 of 148 at B = 96, H = 512); nothing on its stream can
# use the rest.  The serial chain of the encoder's backward is  BPTT(l+1) -> dX GEMM(l+1) -> BPTT(l) -> ...; the dW_ih / dW_hh GEMMs
# and bias sums of a layer feed nothing but the optimizer.  When the layer's parameters keep their gradients in the reducer's flat
# buckets (las_b200.ddp.BucketedGradReducer tags them), backward therefore does not return those gradients through autograd: it
# queues the work, and the NEXT layer's backward issues it on a second stream that is released once every CTA of its own
# BPTT kernel is resident (las_set_launch_start_stream), capped to the SMs that kernel leaves free, accumulating straight into
# p.grad.  What is still queued when the backward pass ends (the bottom layer's own gradients) is issued then, and the calling
# stream waits for the second stream before backward() returns (engine callback), so callers see ordinary semantics.
# LAS_BWD_OVERLAP=0 disables it; parameters without the reducer's tag always take the plain autograd route.
# ----------------------------------------------------------------------------------------------------------------------
class _BackwardOverlap:
    def __init__(self, dev: torch.device):
        self.dev = dev
        self.side = torch.cuda.Stream(device=dev)
        self.pending = []            # closures fn(max_ctas), run with the side stream current
        self.keepalive = []          # main-stream tensors the side stream reads, released after the join
        self.task_id = None          # autograd graph task the queue belongs to
        self.sm_count = torch.cuda.get_device_properties(dev).multi_processor_count

    def run_pending(self, max_ctas: int):
        with torch.cuda.stream(self.side):
            for fn in self.pending:
                fn(max_ctas)
        self.pending = []


_OVERLAP: dict = {}


def _overlap_state(dev: torch.device) -> _BackwardOverlap:
    st = _OVERLAP.get(dev.index)
    if st is None:
        st = _OVERLAP[dev.index] = _BackwardOverlap(dev)
    return st


_DGRAD_STREAMS: dict = {}


def _dgrad_stream(dev: torch.device) -> torch.cuda.Stream:
    """Third stream (high priority: its work is on the critical path of backward, the weight gradients on the second stream are not)."""
    st = _DGRAD_STREAMS.get(dev.index)
    if st is None:
        st = _DGRAD_STREAMS[dev.index] = torch.cuda.Stream(device=dev, priority=-1)
    return st


def _overlap_finish():
    """End of the backward pass (autograd engine callback, caller's thread): issue what is still queued, join the streams."""
    for st in _OVERLAP.values():
        st.task_id = None
        with torch.cuda.device(st.dev):
            main = torch.cuda.current_stream(st.dev)
            if st.pending:
                st.side.wait_stream(main)
                st.run_pending(0)
            main.wait_stream(st.side)
        st.keepalive = []            # main-stream work enqueued from here on is ordered behind everything the side stream read


def _overlap_begin(ovl: _BackwardOverlap) -> None:
    """First deferring node of a backward pass (or leftovers of an aborted one): reset the queue, register the end-of-pass join."""
    tid = torch._C._current_graph_task_id()
    if ovl.task_id != tid:
        ovl.pending = []
        ovl.keepalive = []
        ovl.task_id = tid
        torch.autograd.Variable._execution_engine.queue_callback(_overlap_finish)


def _overlap_ok(wrefs) -> bool:
    # LAS_BWD_OVERLAP=0 switches the overlap off.  It stays on under Nsight Compute: a one-step launch list with the overlap on completes
    # (profiles/launches_r2_step_summary.csv) -- ncu serialises kernels, the cuStreamWaitValue32 hand-off is enqueued after the BPTT
    # launch call returns and finds its counter already past the target; round 1's "did not finish" was a whole bench under ncu.
    if wrefs is None or torch.is_grad_enabled() or os.environ.get('LAS_BWD_OVERLAP', '1') == '0':
        return False
    for w in wrefs:
        g = w.grad
        if (not getattr(w, '_las_bucketed', False)) or (not w.requires_grad) or g is None or g.dtype != torch.float32 or g.shape != w.shape \
                or not g.is_contiguous():
            return False
    return True


def _lstm_weight_grads(dGb, xb, hs, dG, dbp, dims, max_ctas=0, clone_bias=True):
    """[dW_ih, dW_hh, db_ih, db_hh] per direction on the tensor pipe (reference: autograd of nn.LSTM, src/modules.py:80,189).
    dGb (B*T, NG) bf16 gate gradients; xb the layer input as bf16; hs the zero-framed hidden states (bf16, or fp32 to be cast);
    dbp per-slice bias partials from the BPTT kernel (or None: column sums of dG)."""
    Bn, Tin, T, H, ndir, Din, pyramid, Dp, Kp = dims
    F_, G4 = ndir * H, 4 * H
    NG = ndir * G4
    M = Bn * T
    dev = dGb.device
    dwcat = torch.empty(NG, Kp, dtype=torch.float32, device=dev)
    gemm_tc(dGb, xb, dwcat, NG, Kp, T, k_batches=Bn, a_s1=NG, a_s2=T * NG, b_s1=(2 * Dp if pyramid else Dp),
            b_s2=Tin * Dp, ldc=Kp, a_mn=True, b_mn=True, flops=2.0 * Bn * T * NG * Din, max_ctas=max_ctas)
    hsb = hs if hs.dtype == torch.bfloat16 else cast_bf16(hs, Bn * (T + 2), F_, F_, F_)   # (B*(T+2), F) bf16
    grads = []
    for d in range(ndir):
        dw_ih = dwcat[d * G4:(d + 1) * G4, :Din]
        if clone_bias or Kp != Din:
            dw_ih = dw_ih.contiguous()
        dw_hh = torch.empty(G4, H, dtype=torch.float32, device=dev)
        # h_{t-1} for the forward direction is frame t of hs_pad, for the reverse direction frame t+2
        gemm_tc(dGb, hsb, dw_hh, G4, H, T, k_batches=Bn, a_s1=NG, a_s2=T * NG, b_s1=F_, b_s2=(T + 2) * F_, ldc=H,
                a_mn=True, b_mn=True, a_off=d * G4, b_off=(2 * F_ if d == 1 else 0) + d * H, gate=False, max_ctas=max_ctas)
        db = torch.empty(G4, dtype=torch.float32, device=dev)
        if dbp is not None:
            colsum(dbp, G4, dbp.shape[1], G4, db, x_off=d * dbp.shape[1] * G4)      # add the few batch-slice rows
        else:
            colsum(dG, NG, M, G4, db, x_off=d * G4)
        grads += [dw_ih, dw_hh, db, db.clone() if clone_bias else db]
    return grads


# ----------------------------------------------------------------------------------------------------------------------
# One (Bi)LSTM layer over a padded batch with PackedSequence semantics (+ optional pyramidal frame-pair concat and
# locked dropout): reference src/modules.py:74-84 (LockedLSTM loop body) and :165-193 (pyramLockedLSTM loop body).
# ----------------------------------------------------------------------------------------------------------------------
# ----------------------------------------------------------------------------------------------------------------------
# Forward pipelining: the NEXT layer's gate projection beside this layer's recurrence.
#
# The forward recurrence is latency bound and occupies rs * slices * ndir SMs (96 of 148 at B = 96, H = 512).  The next layer's input
# GEMM (reference: inside nn.LSTM, src/modules.py:80 / :189) needs, for output row t, this layer's output at time t (frames 2t, 2t+1
# under the pyramid) from BOTH directions -- available once the forward direction has passed it and the reverse direction has come
# back to it, i.e. from the middle of the sequence outwards.  The recurrence kernel publishes per-cluster progress counters
# (las_lstm_rec_fwd_arm_progress); the GEMM is issued as 128-row time tiles on the second stream, each behind a cuStreamWaitValue32 on
# those counters, in the order the tiles become ready.  The tiles only the last steps release run after the kernel.  The consumer
# (the next lstm_layer call) finds the finished gates on its input tensor (`_las_pregates`) and waits for the second stream's event.
# LAS_FWD_PIPELINE=0 switches it off.
# ----------------------------------------------------------------------------------------------------------------------
_PROGRESS_EVERY = 32
_PIPE_TILE = 128


class _PreGates:
    """Gates of the next layer, projected on the second stream.  Holds every tensor that stream touches: they were allocated on the
    main stream, so their memory may only go back to the allocator once the main stream is ordered behind the second stream's work --
    which the consumer does (wait_event) when it takes the gates over, and __del__ does when nobody did.  (No Tensor.record_stream: its
    event-deferred frees make the caching allocator's steady state depend on timing -- seen as 40 -> 45 ms steps in one run out of three.)"""
    __slots__ = ('gates', 'wcat', 'event', 'x16', 'wkey', 'pyramid', 'T', 'Kp', 'Dp', 'tiles_early', 'tiles_late', 'keep', 'consumed')

    def __del__(self):
        try:
            if not getattr(self, 'consumed', True):
                torch.cuda.current_stream(self.gates.device).wait_event(self.event)
        except Exception:                  # interpreter shutdown
            pass


def _weights_key(weights) -> tuple:
    return tuple((id(w), w._version) for w in weights)


def _pipeline_ok(nxt, x16) -> bool:
    return (nxt is not None and x16 is not None and os.environ.get('LAS_FWD_PIPELINE', '1') != '0'
            and not torch.is_inference_mode_enabled())


def _pipeline_prepare(nxt, Bn, F_, dev):
    """Main stream, BEFORE the recurrence launch: the next layer's bf16 weights / biases / output buffer and the progress counters."""
    ws = [_f32c(w) for w in nxt['weights']]
    ndir = len(ws) // 4
    H = ws[1].shape[1]
    G4 = 4 * H
    NG = ndir * G4
    pyr = bool(nxt['pyramid'])
    Dp = F_
    Kp = 2 * Dp if pyr else Dp
    if Dp % 8 != 0 or Dp < 64 or ws[0].shape[1] != Kp:
        return None
    wcat = torch.empty(NG, Kp, dtype=torch.bfloat16, device=dev)
    for d in range(ndir):
        cast_bf16(ws[4 * d], G4, Kp, Kp, Kp, dst=wcat, dst_off=d * G4 * Kp)
    b1 = torch.cat([ws[4 * d + 2] for d in range(ndir)])
    b2 = torch.cat([ws[4 * d + 3] for d in range(ndir)])
    Tn = int(nxt['T'])
    gates = torch.empty(Bn, Tn, ndir, G4, dtype=torch.float32, device=dev)
    counters = torch.zeros(64, dtype=torch.int32, device=dev)
    return dict(wcat=wcat, b1=b1, b2=b2, gates=gates, counters=counters, NG=NG, Kp=Kp, Dp=Dp, Tn=Tn, pyr=pyr, ndir=ndir, H=H)


def _issue_tiles(stream, counters, ncl, rs, early, late, ev_ready, ev_done, tile_fn):
    """`early` = [(k, t0, t1)] sorted by k: tile_fn(t0, t1) runs on `stream` once every cluster's progress word is >= rs * k;
    `late` tiles run after `ev_done` (the recurrence kernel has finished).  Returns the event recorded behind the last tile."""
    lib = _lib.load()
    cptr = counters.data_ptr()
    with torch.cuda.stream(stream):
        stream.wait_event(ev_ready)
        waited = 0
        for k, t0, t1 in early:
            if k > waited:
                for c in range(ncl):
                    check(lib.las_stream_wait_value_geq(stream.cuda_stream, cptr + 4 * c, rs * k), 'stream_wait_value')
                waited = k
            tile_fn(t0, t1, True)
        stream.wait_event(ev_done)
        for _, t0, t1 in late:
            tile_fn(t0, t1, False)
        ev = torch.cuda.Event()
        ev.record(stream)
    return ev


def _time_tiles(Tn, T, fac, publishes):
    """128-row time tiles [t0, t1) of a (B, Tn) row space whose row t needs recurrence steps < max(fac*t1, T - fac*t0) of BOTH
    directions; (k, t0, t1) with k = the progress count to wait for, None when only the end of the kernel releases the tile."""
    kmax = (T - 1) // _PROGRESS_EVERY
    tiles = []
    for t0 in range(0, Tn, _PIPE_TILE):
        t1 = min(t0 + _PIPE_TILE, Tn)
        ready = max(fac * t1, T - fac * t0)
        k = -(-ready // _PROGRESS_EVERY)
        tiles.append((k if (publishes and k <= kmax) else None, t0, t1))
    return sorted([t for t in tiles if t[0] is not None]), [t for t in tiles if t[0] is None]


def direction_half_schedule(Tn: int, T: int, fac: int, every: int = _PROGRESS_EVERY, tile: int = _PIPE_TILE):
    """Pure host logic of the direction-half pipelining (tests/test_cpu_schedule.py checks it against a step-by-step simulation of the
    two sweeps).  Rows [t0, t1) of the consuming layer read frames [fac*t0, fac*t1) of a T-step BiLSTM layer.  Direction 0 (forward
    sweep, frame s at step s) has written them after min(fac*t1, T) steps, direction 1 (reverse sweep, frame T-1-s at step s) after
    T - fac*t0 steps.  The kernel publishes "steps < every*k are complete" for k = 1 .. (T-1)//every, nothing later.
    Returns (early, late): early = [(k, d, t0, t1)] sorted by k -- half d of the tile may run once direction d's progress count is k;
    late = [(d or None, t0, t1)] -- halves only the end of the kernel releases; d = None: neither half was early, one full-K GEMM."""
    kmax = (T - 1) // every
    early, late = [], []
    for t0 in range(0, Tn, tile):
        t1 = min(t0 + tile, Tn)
        ks = []
        for d in (0, 1):
            ready = min(fac * t1, T) if d == 0 else T - fac * t0
            k = -(-ready // every)
            ks.append(k if k <= kmax else None)
        if ks[0] is None and ks[1] is None:
            late.append((None, t0, t1))
            continue
        for d in (0, 1):
            if ks[d] is not None:
                early.append((ks[d], d, t0, t1))
        for d in (0, 1):
            if ks[d] is None:
                late.append((d, t0, t1))
    early.sort()
    return early, late


def _issue_direction_halves(side, prep, ncl, rs, T, fac, Hp, ev_ready, ev_rec_done, tile_gemm, half_gemm):
    """The same tiles with the reduction split by the producing layer's DIRECTION.  Row t of the next layer's input is
    [h_fwd | h_bwd] of frame t (frames 2t, 2t+1 under the pyramid): the forward sweep has written its half of rows < t1 after fac * t1
    steps, the reverse sweep its half of rows >= t0 after T - fac * t0 steps -- each half of every tile becomes ready at its own time,
    from the first steps on, instead of both sweeps having to cross the tile (the second half of the kernel only).  The half that comes
    first writes (+ biases), the other one accumulates (fp32).  A tile neither half of which is released before the kernel ends runs
    as one full-K GEMM.  Returns (event, halves beside the kernel, GEMMs after it)."""
    lib = _lib.load()
    nper = ncl // 2                                          # word index = direction * nper + batch-slice group (lstm_rec_tc.cu)
    early, late_tiles = direction_half_schedule(prep['Tn'], T, fac)
    started = set()
    cptr = prep['counters'].data_ptr()
    with torch.cuda.stream(side):
        side.wait_event(ev_ready)
        waited = [0, 0]
        for k, d, t0, t1 in early:
            if k > waited[d]:
                for cix in range(d * nper, (d + 1) * nper):
                    check(lib.las_stream_wait_value_geq(side.cuda_stream, cptr + 4 * cix, rs * k), 'stream_wait_value')
                waited[d] = k
            half_gemm(d, t0, t1, t0 not in started, True)
            started.add(t0)
        side.wait_event(ev_rec_done)
        for d, t0, t1 in late_tiles:
            if d is None:
                tile_gemm(t0, t1, False)
            else:
                half_gemm(d, t0, t1, t0 not in started, False)
                started.add(t0)
        ev = torch.cuda.Event()
        ev.record(side)
    return ev, len(early), len(late_tiles)


def _pipeline_issue(prep, nxt, out16, Bn, T, ev_ready, ev_rec_done, ndir_prev=1):
    """After the recurrence launch: the tiles of the next layer's gate GEMM on the second stream, behind the progress counters."""
    lib = _lib.load()
    ncl, rs = C.c_int(0), C.c_int(0)
    publishes = bool(lib.las_lstm_rec_fwd_progress_info(C.byref(ncl), C.byref(rs)))
    dev = out16.device
    side = _overlap_state(dev).side
    fac = 2 if prep['pyr'] else 1
    Tn, NG, Kp, Dp = prep['Tn'], prep['NG'], prep['Kp'], prep['Dp']
    xb = out16.view(Bn * T, -1)
    a_s1 = 2 * Dp if prep['pyr'] else Dp
    gates, wcat, b1, b2 = prep['gates'], prep['wcat'], prep['b1'], prep['b2']

    def tile_gemm(t0, t1, beside):
        R = t1 - t0
        gemm_tc(xb, wcat, gates, R, NG, Kp, a_batches=Bn, a_s1=a_s1, a_s2=T * Dp, b_s1=Kp, c_bs=Tn * NG, ldc=NG,
                bias1=b1, bias2=b2, a_off=t0 * a_s1, c_off=t0 * NG, flops=2.0 * Bn * R * NG * Kp, side=beside)

    Hp = Dp // 2                                             # one direction's width in the producing layer's output
    # LAS_FWD_KSPLIT=0: whole tiles behind BOTH sweeps (bit-identical to the unpipelined GEMM); default: direction halves
    ksplit = (ndir_prev == 2 and publishes and ncl.value >= 2 and ncl.value % 2 == 0 and Dp % 2 == 0 and Hp % 64 == 0
              and os.environ.get('LAS_FWD_KSPLIT', '1') != '0')
    if ksplit:
        def half_gemm(d, t0, t1, first, beside):
            R = t1 - t0
            gemm_tc(xb, wcat, gates, R, NG, fac * Hp, a_batches=Bn, a_s1=a_s1, a_s2=T * Dp, b_s1=Kp, c_bs=Tn * NG, ldc=NG,
                    bias1=b1 if first else None, bias2=b2 if first else None, accumulate=not first, a_off=t0 * a_s1 + d * Hp,
                    b_off=d * Hp, c_off=t0 * NG, flops=2.0 * Bn * R * NG * fac * Hp, side=beside,
                    k_chunk=(Hp if fac == 2 else 0), k_chunk_stride=(Dp if fac == 2 else 0))

        ev, n_early, n_late = _issue_direction_halves(side, prep, ncl.value, rs.value, T, fac, Hp, ev_ready, ev_rec_done, tile_gemm,
                                                      half_gemm)
    else:
        early, late = _time_tiles(Tn, T, fac, publishes)
        ev = _issue_tiles(side, prep['counters'], ncl.value, rs.value, early, late, ev_ready, ev_rec_done, tile_gemm)
        n_early, n_late = len(early), len(late)
    pre = _PreGates()
    pre.gates, pre.wcat, pre.event, pre.x16 = prep['gates'], prep['wcat'], ev, out16
    pre.wkey, pre.pyramid, pre.T, pre.Kp, pre.Dp = _weights_key(nxt['weights']), prep['pyr'], Tn, Kp, Dp
    pre.tiles_early, pre.tiles_late = n_early, n_late
    pre.keep = (prep['counters'], prep['b1'], prep['b2'])        # see the class comment
    pre.consumed = False
    return pre


def _pregates_for(x, x16, pyramid, T, weights):
    """The gates a previous lstm_layer call already projected for THIS layer (same input tensor, unchanged since; same weights)."""
    pre = getattr(x, '_las_pregates', None)
    if pre is None or x16 is None or pre.x16.data_ptr() != x16.data_ptr() or pre.pyramid != bool(pyramid) or pre.T != int(T):
        return None
    if pre.wkey != _weights_key(weights):
        return None
    return pre


_WG_TILE = 256


def _wgrad_tiles_behind_bptt(ovl, wg, dGb, xb, hsb, dbp, wrefs, wdims) -> bool:
    """Weight gradients of the layer that ends the backward pass, as time tiles on the second stream behind the progress counters of
    its own BPTT kernel.  dW of direction d sums over that direction's gate gradients only, and the BPTT sweep of direction 0 produces
    them from the last frame down, that of direction 1 from the first frame up: tiles become ready from the first steps on.  Each
    tile is a K-chunk of the (4H x Din) / (4H x H) GEMMs, accumulated in place (fp32).  Returns False when the kernel that ran does not
    publish progress (the caller then queues the work as before)."""
    lib = _lib.load()
    ncl, rs = C.c_int(0), C.c_int(0)
    if not lib.las_lstm_rec_fwd_progress_info(C.byref(ncl), C.byref(rs)) or ncl.value <= 0:
        return False
    Bn, Tin, T, H, ndir, Din, pyramid, Dp, Kp = wdims
    F_, G4 = ndir * H, 4 * H
    NG = ndir * G4
    dev = dGb.device
    side = ovl.side
    free = ovl.sm_count - 4 * (H // 128) * ((Bn + 31) // 32) * ndir
    if free < 16:
        return False
    nper = ncl.value // ndir                                # clusters (batch-slice groups) per direction; word index = dir * nper + group
    kmax = T // _PROGRESS_EVERY                              # increments a CTA makes: after steps every, 2*every, ... <= T
    main = wg['main']
    ev_done = torch.cuda.Event()
    ev_done.record(main)
    work = []                                                # (k or None, direction, t0, t1)
    for d in range(ndir):
        for t0 in range(0, T, _WG_TILE):
            t1 = min(t0 + _WG_TILE, T)
            ready = (T - t0) if d == 0 else t1               # BPTT steps that must be complete (direction 0 sweeps t = T-1 .. 0)
            k = -(-ready // _PROGRESS_EVERY)
            work.append((k if k <= kmax else None, d, t0, t1))
    early = sorted([wk for wk in work if wk[0] is not None])
    late = [wk for wk in work if wk[0] is None]
    b_s1 = 2 * Dp if pyramid else Dp
    dwcat, dw_hh = wg['dwcat'], wg['dw_hh']

    def tile(d, t0, t1, beside):
        R = t1 - t0
        cap = free if beside else 0
        gemm_tc(dGb, xb, dwcat, G4, Kp, R, k_batches=Bn, a_s1=NG, a_s2=T * NG, b_s1=b_s1, b_s2=Tin * Dp, ldc=Kp, a_mn=True, b_mn=True,
                a_off=d * G4 + t0 * NG, b_off=t0 * b_s1, c_off=d * G4 * Kp, accumulate=True, flops=2.0 * Bn * R * G4 * Din, max_ctas=cap,
                side=beside)
        gemm_tc(dGb, hsb, dw_hh[d], G4, H, R, k_batches=Bn, a_s1=NG, a_s2=T * NG, b_s1=F_, b_s2=(T + 2) * F_, ldc=H, a_mn=True, b_mn=True,
                a_off=d * G4 + t0 * NG, b_off=(2 * F_ if d == 1 else 0) + d * H + t0 * F_, accumulate=True, gate=False, max_ctas=cap)

    cptr = wg['counters'].data_ptr()
    with torch.cuda.stream(side):
        side.wait_event(wg['ev_ready'])
        waited = [0] * ndir
        for k, d, t0, t1 in early:
            if k > waited[d]:
                for cix in range(d * nper, (d + 1) * nper):
                    check(lib.las_stream_wait_value_geq(side.cuda_stream, cptr + 4 * cix, rs.value * k), 'stream_wait_value')
                waited[d] = k
            tile(d, t0, t1, True)
        side.wait_event(ev_done)
        for _, d, t0, t1 in late:
            tile(d, t0, t1, False)
        grads = []
        for d in range(ndir):
            db = torch.empty(G4, dtype=torch.float32, device=dev)
            colsum(dbp, G4, dbp.shape[1], G4, db, x_off=d * dbp.shape[1] * G4)
            grads += [dwcat[d * G4:(d + 1) * G4, :Din], dw_hh[d], db, db]
        for w, g in zip(wrefs, grads):
            w.grad.add_(g)

    def report(max_ctas, wrefs=wrefs):
        # stands in for the post-accumulate-grad hook (bucket all-reduce).  Not now: autograd's own hook call for these parameters (with
        # the None gradients this backward returns) comes AFTER this function and must still find them marked as deferred; the queue
        # runs on the second stream, behind the accumulation above, when the backward pass ends
        for w in wrefs:
            ready = getattr(w, '_las_grad_ready', None)
            if ready is not None:
                ready(w)

    ovl.pending.append(report)
    ovl.keepalive.append((dGb, xb, hsb, dbp, wg['counters'], dwcat, dw_hh))
    last_pipeline_stats[('wgrad', Bn, T, H)] = (len(early), len(late))
    return True


last_pipeline_stats = {}


class LSTMLayerFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, lens_dev, T, pyramid, mask, x16, wrefs, nxt, pre, *weights):
        """x (B, Tin, D) fp32 with contiguous features; lens_dev (B) int32 = lengths AFTER the pyramid halving;
        T = max of those lengths; weights = (w_ih, w_hh, b_ih, b_hh) per direction."""
        _require_cuda(x, lens_dev, *weights)
        lib = _lib.load()
        # the bf16 copy of the output (second return value) feeds only GEMMs outside autograd: without this, backward is handed a
        # zero-filled (B, T, 2H) bf16 "gradient" for it on every layer (0.15 ms of fills per train step in front of the BPTT kernels)
        ctx.set_materialize_grads(False)
        ndir = len(weights) // 4
        assert ndir in (1, 2) and len(weights) == 4 * ndir
        if x.dtype != torch.float32:
            x = x.float()
        if x.stride(2) != 1:
            x = x.contiguous()
        Bn, Tin, D = x.shape
        sb, st = _rows3d(x)
        if pyramid:
            if st != D:            # frame pairs must be adjacent in memory for the concat-as-addressing trick
                x = x.contiguous()
                x16 = None
                sb, st = _rows3d(x)
            Din, st_eff = 2 * D, 2 * st
            assert 2 * T <= Tin
        else:
            Din, st_eff = D, st
            assert T <= Tin
        ws = [_f32c(w) for w in weights]
        H = ws[1].shape[1]
        F_ = ndir * H
        G4 = 4 * H
        NG = ndir * G4
        dev = x.device
        # tensor-pipe mode: bf16 tcgen05 GEMMs for the gate projections (pyramid layers need D % 8 == 0 so that the
        # frame-pair concat stays a pure TMA stride; the 15-dim base layer is zero-padded to K = 64)
        tc = use_tensor_cores() and (D % 8 == 0 if pyramid else True)
        gates = torch.empty(Bn, T, ndir, G4, dtype=torch.float32, device=dev)
        xb = wcat = None
        Dp = Kp = 0
        if tc:
            Dp = D if D % 8 == 0 and D >= 64 else ((D + 63) // 64) * 64
            if pyramid:
                Dp = D
            Kp = 2 * Dp if pyramid else Dp
            if x16 is not None and Dp == D:
                xb = x16.view(Bn * Tin, D)                                                # written by the producing recurrence kernel
            else:
                xb = cast_bf16(x, Bn * Tin, D, Dp, st, inner=Tin, bs=sb)                  # (B*Tin, Dp) bf16, compact
            if pre is not None and pre.Kp == Kp and pre.Dp == Dp and xb.data_ptr() == pre.x16.data_ptr() \
                    and tuple(pre.gates.shape) == tuple(gates.shape):
                # projected beside the previous layer's recurrence (second stream): take it over once that stream is done
                gates, wcat = pre.gates, pre.wcat
                torch.cuda.current_stream(dev).wait_event(pre.event)
                pre.consumed = True
            else:
                wcat = torch.empty(NG, Kp, dtype=torch.bfloat16, device=dev)
                for d in range(ndir):
                    cast_bf16(ws[4 * d], G4, Din, Kp, Din, dst=wcat, dst_off=d * G4 * Kp)
                b1 = torch.cat([ws[4 * d + 2] for d in range(ndir)])
                b2 = torch.cat([ws[4 * d + 3] for d in range(ndir)])
                gemm_tc(xb, wcat, gates, T, NG, Kp, a_batches=Bn, a_s1=(2 * Dp if pyramid else Dp), a_s2=Tin * Dp, b_s1=Kp,
                        c_bs=T * NG, ldc=NG, bias1=b1, bias2=b2, lens=lens_dev, flops=2.0 * Bn * T * NG * Din)
        else:
            for d in range(ndir):
                w_ih, _, b_ih, b_hh = ws[4 * d:4 * d + 4]
                assert w_ih.shape == (G4, Din), (w_ih.shape, G4, Din)
                gemm_raw(x, w_ih, gates, Bn * T, G4, Din, am=(sb, st_eff, T), ak=(0, 1, 0), bk=(0, 1, 0), bn=Din,
                         cm=(0, NG, 0), bias1=b_ih, bias2=b_hh, c_off=d * G4, gate=True)
        w_hh = torch.stack([ws[4 * d + 1] for d in range(ndir)], 0).contiguous()
        rec_tc = tc and os.environ.get('LAS_REC_TC', '1') == '1' and bool(lib.las_lstm_rec_tc_supported(Bn, H, ndir))
        train = any(ctx.needs_input_grad)
        # tensor-pipe recurrence: what only GEMMs read afterwards (the next layer's input, the dW_hh operand) is written as bf16 by
        # the kernel itself; the fp32 hidden states are then only needed when they ARE the output (no locked-dropout mask)
        shadows = rec_tc and os.environ.get('LAS_REC_BF16_OUT', '1') == '1'
        out16 = torch.empty(Bn, T, F_, dtype=torch.bfloat16, device=dev) if shadows else None
        hs16 = torch.empty(Bn, T + 2, F_, dtype=torch.bfloat16, device=dev) if (shadows and train) else None
        need_hs32 = (mask is None) or hs16 is None
        hs_pad = torch.empty(Bn, T + 2, F_, dtype=torch.float32, device=dev) if need_hs32 else None
        cs_pad = torch.empty(Bn, T + 2, F_, dtype=torch.float32, device=dev)
        out = torch.empty(Bn, T, F_, dtype=torch.float32, device=dev) if mask is not None else None
        if mask is not None:
            mask = _f32c(mask).reshape(Bn, F_)
        pre_next = None
        if rec_tc:
            # tensor-pipe recurrence: W_hh as bf16 (ndir*4H, H), resident in shared memory inside the kernel
            w_hh_b = cast_bf16(w_hh, ndir * G4, H, H, H)
            nbytes = lib.las_lstm_rec_tc_workspace_bytes(Bn, H, ndir)
            wsb = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            prep = _pipeline_prepare(nxt, Bn, F_, dev) if _pipeline_ok(nxt, out16) else None
            if prep is not None:
                main = torch.cuda.current_stream(dev)
                ev_ready = torch.cuda.Event()
                ev_ready.record(main)
                lib.las_lstm_rec_fwd_arm_progress(prep['counters'].data_ptr(), _PROGRESS_EVERY)
            check(lib.las_lstm_rec_fwd_tc_ex(gates.data_ptr(), w_hh_b.data_ptr(), lens_dev.data_ptr(), ptr(mask), ptr(out),
                                             ptr(hs_pad), cs_pad.data_ptr(), Bn, T, H, ndir, int(train), wsb.data_ptr(),
                                             nbytes, ptr(out16), ptr(hs16), stream_ptr()), 'lstm_rec_fwd_tc')
            if prep is not None:
                ev_rec_done = torch.cuda.Event()
                ev_rec_done.record(main)
                pre_next = _pipeline_issue(prep, nxt, out16, Bn, T, ev_ready, ev_rec_done, ndir_prev=ndir)
                last_pipeline_stats[(Bn, T, H)] = (pre_next.tiles_early, pre_next.tiles_late)
        else:
            nbytes = lib.las_lstm_rec_workspace_bytes(Bn, H, ndir)
            wsb = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            check(lib.las_lstm_rec_fwd_f32(gates.data_ptr(), w_hh.data_ptr(), lens_dev.data_ptr(), ptr(mask), ptr(out),
                                           hs_pad.data_ptr(), cs_pad.data_ptr(), Bn, T, H, ndir, wsb.data_ptr(), nbytes,
                                           stream_ptr()), 'lstm_rec_fwd')
        if tc:
            ctx.save_for_backward(xb, lens_dev, gates, hs16 if hs16 is not None else hs_pad, cs_pad, w_hh, mask, wcat)
        else:
            ctx.save_for_backward(x, lens_dev, gates, hs_pad, cs_pad, w_hh, mask, *ws)
        ctx.dims = (Bn, Tin, D, T, H, ndir, Din, sb, st_eff, bool(pyramid), tc, Dp, Kp)
        ctx.wrefs = wrefs if (tc and rec_tc) else None
        y = out if out is not None else hs_pad[:, 1:T + 1]
        if out16 is not None:
            ctx.mark_non_differentiable(out16)
        ctx.pre_next = None
        LSTMLayerFunction._handoff = pre_next          # picked up by lstm_layer right after apply() (not a tensor: cannot be returned)
        return y, out16

    @staticmethod
    def backward(ctx, dy, _d16=None):
        lib = _lib.load()
        x, lens_dev, gates, hs_pad, cs_pad, w_hh, mask, *ws = ctx.saved_tensors
        Bn, Tin, D, T, H, ndir, Din, sb, st_eff, pyramid, tc, Dp, Kp = ctx.dims
        F_, G4 = ndir * H, 4 * H
        NG = ndir * G4
        dev = x.device
        if dy is None:                     # the layer's output was not used (gradients are not materialised, see forward)
            dy = torch.zeros(Bn, T, ndir * H, dtype=torch.float32, device=dev)
        dy = _f32c(dy)
        rec_tc = tc and os.environ.get('LAS_REC_TC', '1') == '1' and bool(lib.las_lstm_rec_tc_supported(Bn, H, ndir))
        nbytes = lib.las_lstm_rec_tc_workspace_bytes(Bn, H, ndir) if rec_tc else lib.las_lstm_rec_workspace_bytes(Bn, H, ndir)
        wsb = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        # gates (activated) -> d(pre-activation), in place.  The forward's saved tensor is consumed: a second
        # backward through the same graph is not supported (like cuDNN's reserve space, it is single use).
        M = Bn * T
        dGb = None
        dbp = None
        ovl = _overlap_state(dev) if (rec_tc and _overlap_ok(ctx.wrefs)) else None
        ev = False
        if ovl is not None:
            _overlap_begin(ovl)
            if ovl.pending:
                ev = True
                lib.las_set_launch_start_stream(ovl.side.cuda_stream)
        dx_pipe = wg_pipe = None
        if rec_tc:
            w_hh_t = torch.empty(ndir, H, G4, dtype=torch.bfloat16, device=dev)
            check(lib.las_transpose_cast_bf16(w_hh.data_ptr(), w_hh_t.data_ptr(), ndir, G4, H, stream_ptr()), 'transpose_cast')
            dGb = torch.empty(M, NG, dtype=torch.bfloat16, device=dev)
            want_dx = tc and ctx.needs_input_grad[0] and os.environ.get('LAS_BWD_PIPELINE', '1') != '0' and T > _PIPE_TILE
            # weight gradients behind the layer's OWN BPTT kernel (direction by direction, from its first steps on): for the layer that
            # ends the backward pass nothing else could hide them; for a long layer above it this empties the next BPTT kernel's window,
            # which the shorter kernels of round 2 no longer covered
            wg_mode = os.environ.get('LAS_BWD_WGRAD_PIPELINE', '1')          # 0: off, 1: the layer that ends the backward pass, 2: every long layer
            want_wg = tc and ovl is not None and wg_mode != '0' and T > 2 * _WG_TILE and hs_pad.dtype == torch.bfloat16 \
                and (wg_mode == '2' or not ctx.needs_input_grad[0])
            if want_dx or want_wg:
                main = torch.cuda.current_stream(dev)
                counters = torch.zeros(64, dtype=torch.int32, device=dev)        # one set of progress words serves both consumers
                if want_dx:
                    # the layer's dX GEMM, tile by tile beside its own BPTT kernel (row t has both directions' gate gradients once the
                    # two sweeps have crossed it: from the middle outwards), on a third stream: dX is what the layer below waits for
                    # the tiles cover every row t < T of every utterance (no length skipping on this route): only the frames the layer
                    # never read (the odd frame the pyramid drops, src/modules.py:171-175) need zeros -- not a fill of the whole
                    # (B, Tin, D) gradient in front of the BPTT kernel (0.15 ms per train step)
                    dx_new = torch.empty(Bn, Tin, D, dtype=torch.float32, device=dev)
                    used = (T * Din) // D
                    if used < Tin:
                        dx_new[:, used:].zero_()
                    dx_pipe = dict(dx=dx_new, counters=counters, ev_ready=torch.cuda.Event(), main=main)
                if want_wg:
                    wg_pipe = dict(counters=counters, ev_ready=torch.cuda.Event(), main=main,
                                   dwcat=torch.zeros(NG, Kp, dtype=torch.float32, device=dev),
                                   dw_hh=[torch.zeros(G4, H, dtype=torch.float32, device=dev) for _ in range(ndir)])
                for pipe in (dx_pipe, wg_pipe):
                    if pipe is not None:
                        pipe['ev_ready'].record(main)
                lib.las_lstm_rec_bwd_arm_progress(counters.data_ptr(), _PROGRESS_EVERY)
            nsl = lib.las_lstm_rec_bwd_tc_dbias_slices(Bn, H, ndir) if os.environ.get('LAS_REC_DBIAS', '1') == '1' else 0
            if nsl > 0:
                # bias gradients accumulated inside the BPTT kernel (per direction and batch slice); no fp32 dG write-back, no
                # column-sum pass over (B*T, 8H)
                dbp = torch.empty(ndir, nsl, G4, dtype=torch.float32, device=dev)
                check(lib.las_lstm_rec_bwd_tc_db(dy.data_ptr(), gates.data_ptr(), dGb.data_ptr(), cs_pad.data_ptr(), w_hh_t.data_ptr(),
                                                 lens_dev.data_ptr(), ptr(mask), Bn, T, H, ndir, wsb.data_ptr(), nbytes, dbp.data_ptr(),
                                                 stream_ptr()), 'lstm_rec_bwd_tc_db')
            else:
                check(lib.las_lstm_rec_bwd_tc(dy.data_ptr(), gates.data_ptr(), dGb.data_ptr(), cs_pad.data_ptr(), w_hh_t.data_ptr(),
                                              lens_dev.data_ptr(), ptr(mask), Bn, T, H, ndir, wsb.data_ptr(), nbytes, stream_ptr()),
                      'lstm_rec_bwd_tc')
        else:
            check(lib.las_lstm_rec_bwd_f32(dy.data_ptr(), gates.data_ptr(), cs_pad.data_ptr(), w_hh.data_ptr(), lens_dev.data_ptr(),
                                           ptr(mask), Bn, T, H, ndir, wsb.data_ptr(), nbytes, stream_ptr()), 'lstm_rec_bwd')
        if ev:
            # the layer above left its weight-gradient work queued: it runs beside this layer's BPTT kernel, on the SMs that
            # kernel does not occupy (one CTA per (gate-row slice, batch slice, direction), lstm_rec_tc.cu)
            at_start = bool(lib.las_launch_start_mode())      # else the side stream was released behind the kernel, not beside it
            lib.las_set_launch_start_stream(None)
            free = ovl.sm_count - 4 * (H // 128) * ((Bn + 31) // 32) * ndir
            ovl.run_pending(free if (at_start and free >= 16) else 0)
        dG = gates
        dx = None
        grads: List[Optional[torch.Tensor]] = []
        if tc:
            xb, wcat = x, ws[0]
            if dGb is None:
                dGb = cast_bf16(dG, M, NG, NG, NG)                                        # (B*T, NG) bf16
            if dx_pipe is not None:
                ncl, rs = C.c_int(0), C.c_int(0)
                publishes = bool(lib.las_lstm_rec_fwd_progress_info(C.byref(ncl), C.byref(rs)))
                main = dx_pipe['main']
                ev_done = torch.cuda.Event()
                ev_done.record(main)
                dx = dx_pipe['dx']
                early, late = _time_tiles(T, T, 1, publishes)

                def dx_tile(t0, t1, beside=False, dGb=dGb, wcat=wcat, dx=dx):
                    gemm_tc(dGb, wcat, dx, t1 - t0, Din, NG, a_batches=Bn, a_s1=NG, a_s2=T * NG, b_s1=Kp, b_mn=True, c_bs=Tin * D,
                            ldc=Din, a_off=t0 * NG, c_off=t0 * Din, flops=2.0 * Bn * (t1 - t0) * NG * Din, side=beside)

                if early:
                    third = _dgrad_stream(dev)
                    ev = _issue_tiles(third, dx_pipe['counters'], ncl.value, rs.value, early, late, dx_pipe['ev_ready'], ev_done, dx_tile)
                    main.wait_event(ev)                  # everything the main stream does from here on is ordered behind the tiles
                else:
                    for _, t0, t1 in late:
                        dx_tile(t0, t1)
                last_pipeline_stats[('bwd', Bn, T, H)] = (len(early), len(late))
            elif ctx.needs_input_grad[0]:
                dx = torch.zeros(Bn, Tin, D, dtype=torch.float32, device=dev)             # tiles past a row's length are skipped
                gemm_tc(dGb, wcat, dx, T, Din, NG, a_batches=Bn, a_s1=NG, a_s2=T * NG, b_s1=Kp, b_mn=True, c_bs=Tin * D,
                        ldc=Din, lens=lens_dev, flops=2.0 * Bn * T * NG * Din)
            wdims = (Bn, Tin, T, H, ndir, Din, pyramid, Dp, Kp)
            if ovl is not None:
                wrefs = ctx.wrefs

                def run(max_ctas, dGb=dGb, xb=xb, hs=hs_pad, dG=dG, dbp=dbp, wrefs=wrefs, wdims=wdims):
                    for w, g in zip(wrefs, _lstm_weight_grads(dGb, xb, hs, dG, dbp, wdims, max_ctas=max_ctas, clone_bias=False)):
                        w.grad.add_(g)
                    for w in wrefs:                        # stands in for the post-accumulate-grad hook (bucket all-reduce)
                        ready = getattr(w, '_las_grad_ready', None)
                        if ready is not None:
                            ready(w)
                    # allocated on the main stream, read here: kept alive until the streams have joined (no record_stream: its
                    # event-deferred frees make the caching allocator's steady state depend on timing)
                    _overlap_state(dGb.device).keepalive.append((dGb, xb, hs, dG, dbp))

                for w in wrefs:                            # the reducer's autograd hook must not count the None returned below
                    w._las_deferred = True
                if wg_pipe is not None and dbp is not None and _wgrad_tiles_behind_bptt(ovl, wg_pipe, dGb, xb, hs_pad, dbp, wrefs, wdims):
                    return (dx, None, None, None, None, None, None, None, None, *([None] * (4 * ndir)))
                ovl.pending.append(run)
                return (dx, None, None, None, None, None, None, None, None, *([None] * (4 * ndir)))
            grads = _lstm_weight_grads(dGb, xb, hs_pad, dG, dbp, wdims)
            return (dx, None, None, None, None, None, None, None, None, *grads)
        if ctx.needs_input_grad[0]:
            full = (Tin * D == T * Din)
            dx = (torch.empty if full else torch.zeros)(Bn, Tin, D, dtype=torch.float32, device=dev)
            for d in range(ndir):
                gemm_raw(dG, ws[4 * d], dx, M, Din, G4, am=(0, NG, 0), ak=(0, 1, 0), bk=(0, Din, 0), bn=1,
                         cm=(Tin * D, Din, T), beta=0.0 if d == 0 else 1.0, a_off=d * G4, gate=True)
        for d in range(ndir):
            dw_ih = torch.empty(G4, Din, dtype=torch.float32, device=dev)
            gemm_raw(dG, x, dw_ih, G4, Din, M, am=(0, 1, 0), ak=(0, NG, 0), bk=(sb, st_eff, T), bn=1, cm=(0, Din, 0),
                     a_off=d * G4, gate=True)
            dw_hh = torch.empty(G4, H, dtype=torch.float32, device=dev)
            # h_{t-1} for the forward direction is frame t of hs_pad, for the reverse direction frame t+2
            gemm_raw(dG, hs_pad, dw_hh, G4, H, M, am=(0, 1, 0), ak=(0, NG, 0), bk=((T + 2) * F_, F_, T), bn=1,
                     cm=(0, H, 0), a_off=d * G4, b_off=(2 * F_ if d == 1 else 0) + d * H)
            db = torch.empty(G4, dtype=torch.float32, device=dev)
            colsum(dG, NG, M, G4, db, x_off=d * G4)
            grads += [dw_ih, dw_hh, db, db.clone()]
        return (dx, None, None, None, None, None, None, None, None, *grads)


def lstm_layer(x, lens_dev, T, pyramid, mask, weights: Sequence[torch.Tensor], next_layer: Optional[dict] = None):
    """next_layer (optional): dict(weights=<the following layer's 4*ndir tensors>, pyramid=bool, T=<its sequence length>) -- lets this
    layer project the following layer's gates beside its own recurrence (see "Forward pipelining" above)."""
    x16 = _bf16_shadow(x)
    pre = _pregates_for(x, x16, pyramid, T, weights)
    LSTMLayerFunction._handoff = None
    y, y16 = LSTMLayerFunction.apply(x, lens_dev, int(T), bool(pyramid), mask, x16, tuple(weights), next_layer, pre, *weights)
    _attach_bf16_shadow(y, y16)
    handoff, LSTMLayerFunction._handoff = LSTMLayerFunction._handoff, None
    if handoff is not None and y16 is not None and getattr(y, '_las_bf16', None) is not None:
        y._las_pregates = handoff
    return y


# ----------------------------------------------------------------------------------------------------------------------
# Fused attention step (reference src/models.py:168-185), standalone autograd form used by
# MultiheadCrossAttention.forward; the Speller's fused loop calls the same kernels from C.
# ----------------------------------------------------------------------------------------------------------------------
class AttnStepFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, K, V, lens_dev, heads, fmask=None):
        _require_cuda(q, K, V, lens_dev)
        q, K, V = _f32c(q), _f32c(K), _f32c(V)
        Bn, T, P = K.shape
        ctxv = torch.empty(Bn, P, dtype=torch.float32, device=q.device)
        w = torch.empty(Bn, heads, T, dtype=torch.float32, device=q.device)
        d = LasAttnStep()
        w2 = None
        if fmask is not None:         # init-force prior, (B*heads, T) contiguous
            _require_cuda(fmask)
            fmask = _f32c(fmask)
            w2 = torch.empty_like(w)
            d.fmask, d.ld_fmask, d.w2 = fmask.data_ptr(), T, w2.data_ptr()
        d.q, d.ld_q = q.data_ptr(), P
        d.K, d.V, d.lens = K.data_ptr(), V.data_ptr(), lens_dev.data_ptr()
        d.w, d.ld_w = w.data_ptr(), T
        d.ctx, d.ld_ctx = ctxv.data_ptr(), P
        d.B, d.T, d.P, d.heads = Bn, T, P, int(heads)
        d.scale = float((P // heads) ** 0.5)
        check(_lib.load().las_attn_step_fwd_f32(C.byref(d), stream_ptr()), 'attn_step_fwd')
        ctx.save_for_backward(q, K, V, lens_dev, w, ctxv, fmask, w2)
        ctx.heads = int(heads)
        ctx.mark_non_differentiable(w)
        return ctxv, w

    @staticmethod
    def backward(ctx, dctx, _dw):
        q, K, V, lens_dev, w, ctxv, fmask, w2 = ctx.saved_tensors
        heads = ctx.heads
        Bn, T, P = K.shape
        dh = P // heads
        dctx = _f32c(dctx).clone()
        dq = torch.empty_like(q)
        de = torch.empty_like(w)
        d = LasAttnStep()
        d.q, d.ld_q = q.data_ptr(), P
        d.K, d.V, d.lens = K.data_ptr(), V.data_ptr(), lens_dev.data_ptr()
        d.w, d.ld_w = w.data_ptr(), T
        d.ctx, d.ld_ctx = ctxv.data_ptr(), P      # saved context: lets backward run as one pass (dot = dctx . ctx)
        d.dctx, d.ld_dctx = dctx.data_ptr(), P
        d.dq, d.ld_dq, d.dq_accumulate = dq.data_ptr(), P, 0
        d.de = de.data_ptr()
        d.B, d.T, d.P, d.heads = Bn, T, P, heads
        d.scale = float(dh ** 0.5)
        if fmask is not None:
            d.fmask, d.ld_fmask, d.w2 = fmask.data_ptr(), T, w2.data_ptr()
        wv = w if fmask is None else w2           # the weights that multiplied V
        check(_lib.load().las_attn_step_bwd_f32(C.byref(d), stream_ptr()), 'attn_step_bwd')
        dK = dV = None
        if ctx.needs_input_grad[1]:
            dK = torch.empty_like(K)
            for h in range(heads):     # dK[b,t,h*d+j] = de[b,h,t] * q[b,h*d+j]
                gemm_raw(de, q, dK, T, dh, 1, am=(0, 1, 0), ak=(0, 0, 0), bk=(0, 0, 0), bn=1, cm=(0, P, 0), batch=Bn,
                         bsA=heads * T, bsB=P, bsC=T * P, a_off=h * T, b_off=h * dh, c_off=h * dh)
        if ctx.needs_input_grad[2]:
            dV = torch.empty_like(V)
            for h in range(heads):
                gemm_raw(wv, dctx, dV, T, dh, 1, am=(0, 1, 0), ak=(0, 0, 0), bk=(0, 0, 0), bn=1, cm=(0, P, 0), batch=Bn,
                         bsA=heads * T, bsB=P, bsC=T * P, a_off=h * T, b_off=h * dh, c_off=h * dh)
        return dq, dK, dV, None, None, None


def attn_step(q, K, V, lens_dev, heads, fmask=None):
    """fmask: optional (B*heads, T) init-force prior (reference src/models.py:177-181); returns (ctx, pre-prior weights)."""
    return AttnStepFunction.apply(q, K, V, lens_dev, heads, fmask)


# ----------------------------------------------------------------------------------------------------------------------
# Speller decoder loop (reference src/models.py:336-385)
# ----------------------------------------------------------------------------------------------------------------------
SPELLER_PARAM_ORDER = ('emb', 'cls_b', 'w_ih0', 'w_hh0', 'b_ih0', 'b_hh0', 'w_ih1', 'w_hh1', 'b_ih1', 'b_hh1', 'wq', 'bq',
                       'init_query')


def _speller_desc(K, V, enc_lens, params, dec_y, use_gold, drop0, drop1, steps, heads, sos_idx, pad_idx, training, use_tc=False,
                  init_force=False):
    Bn, T, P = K.shape
    emb = params[0]
    s = LasSpeller()
    s.B, s.T, s.P = Bn, T, P
    s.E = emb.shape[1]
    s.DH = params[3].shape[1]
    s.DO = params[7].shape[1]
    s.V = emb.shape[0]
    s.heads, s.steps = int(heads), int(steps)
    s.sos_idx, s.pad_idx = int(sos_idx), int(pad_idx)
    s.training = int(training)
    s.init_force = int(bool(init_force))
    s.use_tc = int(use_tc and s.P % 8 == 0 and s.DH % 8 == 0 and s.DO % 8 == 0)
    for name, t in zip(SPELLER_PARAM_ORDER, params):
        setattr(s, name, t.data_ptr())
    s.K, s.V_, s.enc_lens = K.data_ptr(), V.data_ptr(), enc_lens.data_ptr()
    if dec_y is not None:
        s.dec_y, s.ld_y = dec_y.data_ptr(), dec_y.stride(0)
    keep = None
    if use_gold is not None:
        keep = (C.c_ubyte * len(use_gold))(*[1 if u else 0 for u in use_gold])
        s.use_gold_host = C.cast(keep, C.c_void_p)
    s.drop0, s.drop1 = ptr(drop0), ptr(drop1)
    return s, keep


class LazyHostTensor(torch.Tensor):
    """CPU tensor whose bytes are still on their way from the device.

    The reference returns the attention map of sample 0 as a CPU tensor from every forward (src/models.py:349,377,385); a
    blocking copy there stalls the host until the decoder loop has drained, so the loss and the whole backward are then
    enqueued with no lead over the GPU (~0.7 ms of idle GPU per batch).  This subclass owns a pinned-host copy issued with
    `non_blocking=True` plus the CUDA event recorded after it; the first torch operation / method / property that touches the
    tensor (`.numpy()`, indexing, printing, `np.asarray`, pickling ...) waits for the event, then behaves like the plain tensor.
    """

    @staticmethod
    def __new__(cls, data, event):
        t = torch.Tensor._make_subclass(cls, data, False)
        t._las_event = event
        return t

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        def plain(a):
            if isinstance(a, LazyHostTensor):
                ev = a.__dict__.get('_las_event')
                if ev is not None:
                    ev.synchronize()
                    a.__dict__['_las_event'] = None
                return a.as_subclass(torch.Tensor)
            if isinstance(a, (list, tuple)):
                return type(a)(plain(x) for x in a)
            return a
        with torch._C.DisableTorchFunctionSubclass():
            return func(*[plain(a) for a in args], **{k: plain(v) for k, v in (kwargs or {}).items()})


def host_copy_lazy(t: torch.Tensor) -> torch.Tensor:
    """`t.cpu()` without the host wait (see LazyHostTensor).  LAS_ATT_SYNC=1 -> the plain blocking copy."""
    if (not t.is_cuda) or os.environ.get('LAS_ATT_SYNC', '0') == '1':
        return t.cpu()
    host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    host.copy_(t, non_blocking=True)
    ev = torch.cuda.Event()
    ev.record(torch.cuda.current_stream(t.device))
    return LazyHostTensor(host, ev)


class _SpellerSlot:
    """One pointer-stable buffer set for a decoder loop: the staged inputs (K, V, lengths, gold tokens, dropout masks), the
    workspaces and the raw outputs.  The launch-per-stage loop of las_speller_fwd_f32 replays a CUDA graph keyed on the descriptor
    (every pointer in it), so a steady training / decoding loop must present the SAME pointers every step -- PyTorch's caching
    allocator does not promise that for per-step allocations, a slot does.

    Buffers are flat, grow-only and shared by every shape that fits: a ragged-batch epoch (a new (T, steps) almost every batch, the
    reference trainer's normal case) settles on ONE buffer set sized for its largest batch instead of one set per distinct shape."""
    __slots__ = ('t', 'v', 'busy')

    def __init__(self):
        self.t, self.v, self.busy = {}, {}, False

    def buf(self, name, shape, dtype, device):
        n = 1
        for d in shape:
            n *= int(d)
        b = self.t.get(name)
        if b is None or b.dtype != dtype or b.device != device or b.numel() < n:
            cap = n if b is None or b.dtype != dtype or b.device != device else max(n, int(b.numel() * 1.25))
            cap = max(1, (cap + 255) // 256 * 256)
            # never an inference tensor: the reference evaluates under torch.inference_mode() (src/train.py:207) and decodes test data
            # outside it (src/infer.py:56-62); a buffer born inside could not be written in place afterwards
            with torch.inference_mode(False):
                b = torch.empty(cap, dtype=dtype, device=device)
            self.t[name] = b
        view = b[:n].view(*shape)
        self.v[name] = view
        return view

    def nbytes(self):
        return sum(b.numel() * b.element_size() for b in self.t.values())


class _SlotLease:
    """Held by the autograd node of a training forward; frees the slot when backward has run or the graph is dropped."""
    __slots__ = ('slot',)

    def __init__(self, slot):
        self.slot = slot
        slot.busy = True

    def release(self):
        if self.slot is not None:
            self.slot.busy = False
            self.slot = None

    def __del__(self):
        self.release()


_SPELLER_POOL = {}        # device index -> list of slots (one per concurrently live forward, plus at most _POOL_SPARE idle ones)
_POOL_SPARE = 2


def _acquire_slot(key):
    """key = (device index, training): eval slots carry no history, so they are kept apart from the (much larger) training slots."""
    slots = _SPELLER_POOL.setdefault(key, [])
    idle = [sl for sl in slots if not sl.busy]
    if idle:
        # largest first: the slot most likely to fit without growing; surplus idle slots (left by a burst of live forwards) are dropped
        idle.sort(key=lambda sl: -sl.nbytes())
        for extra in idle[_POOL_SPARE:]:
            slots.remove(extra)
        return idle[0]
    sl = _SpellerSlot()
    slots.append(sl)
    return sl


def speller_pool_clear():
    """Drop every pooled decoder buffer set (they are retained between steps on purpose)."""
    _SPELLER_POOL.clear()


def speller_pool_bytes():
    """Device bytes currently held by the pooled decoder buffer sets."""
    return sum(sl.nbytes() for slots in _SPELLER_POOL.values() for sl in slots)


class SpellerFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, K, V, enc_lens, dec_y, use_gold, drop0, drop1, steps, heads, sos_idx, pad_idx, training, init_force, wrefs, *params):
        _require_cuda(K, V, enc_lens, *params)
        lib = _lib.load()
        ctx.wrefs = wrefs                # the caller's Parameter objects when it allows their gradients to be accumulated out of band
        params = tuple(_f32c(p) for p in params)
        dev = K.device
        Bn, T, P = K.shape
        Vn = params[0].shape[0]
        use_tc = use_tensor_cores() and os.environ.get('LAS_DEC_TC', '1') == '1'
        # the persistent decoder-step kernel (one launch per loop) takes fp32 or fp16 K / V; the launch-per-stage loop fp32 or bf16
        persist = bool(lib.las_speller_persistent(Bn, T, P, params[3].shape[1], params[7].shape[1], Vn, int(heads), int(bool(init_force)),
                                                  int(bool(use_tc))))
        if persist:
            # fp16 K / V rows: both attention passes run on the tensor pipe (mma.sync, fp32 accumulate) and stream half the bytes:
            # 24.3 -> 21.3 us per decoder step at the train shape, 99.9 -> 90.1 ms per greedy batch; logits 9.4e-4 -> 1.0e-3 from the fp32
            # reference at T = 1600, L = 300 (the reference under autocast computes K and V in 16 bits as well).  LAS_KV_F16=0: fp32 rows
            kv16 = 2 if os.environ.get('LAS_KV_F16', '1') == '1' else 0
        else:
            kv16 = 1 if (bool(use_tc) and P % 4 == 0 and os.environ.get('LAS_KV_BF16', '0') == '1') else 0   # measured slower than fp32 rows: off
        slot = _acquire_slot((dev.index, bool(training)))
        # ---- stage the per-step inputs into the slot (device-to-device copies, a few tens of microseconds) ----
        if kv16:
            # AMP mode option: K and V as 16-bit rows (half the bytes per step); the energies and the context still accumulate in fp32
            kdt = torch.float16 if kv16 == 2 else torch.bfloat16
            Kb = slot.buf('K', (Bn, T, P), kdt, dev)
            Vb = slot.buf('V', (Bn, T, P), kdt, dev)
            Kb.copy_(K); Vb.copy_(V)
            if training and kv16 == 2:
                # the backward loop's attention kernels read fp32 (or bf16) rows: keep an fp32 copy beside the fp16 one
                slot.buf('K32', (Bn, T, P), torch.float32, dev).copy_(K)
                slot.buf('V32', (Bn, T, P), torch.float32, dev).copy_(V)
        else:
            Kb = slot.buf('K', (Bn, T, P), torch.float32, dev)
            Vb = slot.buf('V', (Bn, T, P), torch.float32, dev)
            Kb.copy_(K); Vb.copy_(V)
        lens_b = slot.buf('lens', (Bn,), torch.int32, dev)
        lens_b.copy_(enc_lens)
        y_b = None
        if dec_y is not None:
            y_b = slot.buf('y', tuple(dec_y.shape), torch.int32, dev)
            y_b.copy_(dec_y)
        d0 = d1 = None
        if drop0 is not None:
            d0 = slot.buf('drop0', tuple(drop0.shape), torch.float32, dev)
            d1 = slot.buf('drop1', tuple(drop1.shape), torch.float32, dev)
            d0.copy_(drop0); d1.copy_(drop1)
        s, keep = _speller_desc(Kb, Vb, lens_b, params, y_b, use_gold, d0, d1, steps, heads, sos_idx, pad_idx, training, use_tc,
                                init_force)
        s.kv_bf16 = int(kv16)
        nf = lib.las_speller_workspace_floats(C.byref(s))
        ni = lib.las_speller_workspace_ints(C.byref(s))
        fws = slot.buf('fws', (nf,), torch.float32, dev)
        iws = slot.buf('iws', (ni,), torch.int32, dev)
        logits_b = slot.buf('logits', (Bn, steps, Vn), torch.float32, dev)
        att0_b = slot.buf('att0', (steps + 1, heads, T), torch.float32, dev)
        chars_b = slot.buf('chars', (steps, Bn), torch.int32, dev)
        chars_b.zero_()         # steps that do not compute per-step logits (full teacher forcing) leave their row at 0
        s.logits, s.att0, s.chars = logits_b.data_ptr(), att0_b.data_ptr(), chars_b.data_ptr()
        s.fws, s.fws_floats, s.iws, s.iws_ints = fws.data_ptr(), nf, iws.data_ptr(), ni
        check(lib.las_speller_fwd_f32(C.byref(s), stream_ptr()), 'speller_fwd')
        logits, att0, chars = logits_b.clone(), att0_b.clone(), chars_b.clone()      # the caller owns its outputs
        if training:
            ctx.lease = _SlotLease(slot)
            ctx.views = dict(slot.v)
            if drop0 is None:
                ctx.views.pop('drop0', None); ctx.views.pop('drop1', None)
            ctx.save_for_backward(*params)
            ctx.cfg = (use_gold, steps, heads, sos_idx, pad_idx, use_tc, kv16, init_force)
        ctx.mark_non_differentiable(att0, chars)
        return logits, att0, chars

    @staticmethod
    def backward(ctx, dlogits, _datt, _dchars):
        lib = _lib.load()
        params = ctx.saved_tensors
        lease = ctx.lease
        if lease.slot is None:
            raise RuntimeError('speller backward called twice: the decoder history buffers were already released')
        t = ctx.views                     # this forward's views into the slot's buffers
        K, V, enc_lens, dec_y, drop0, drop1, fws, iws = t['K'], t['V'], t['lens'], t.get('y'), t.get('drop0'), t.get('drop1'), t['fws'], t['iws']
        use_gold, steps, heads, sos_idx, pad_idx, use_tc, kv16, init_force = ctx.cfg
        K16 = V16 = None
        if kv16 == 2:
            K16, V16 = K, V                # the forward's fp16 rows: the backward attention step streams them through the tensor pipe
            K, V, kv16 = t['K32'], t['V32'], 0
        s, keep = _speller_desc(K, V, enc_lens, params, dec_y, use_gold, drop0, drop1, steps, heads, sos_idx, pad_idx, True, use_tc,
                                init_force)
        s.kv_bf16 = int(kv16)
        if K16 is not None:
            s.K_f16, s.V_f16 = K16.data_ptr(), V16.data_ptr()
        s.fws, s.fws_floats, s.iws, s.iws_ints = fws.data_ptr(), fws.numel(), iws.data_ptr(), iws.numel()
        # outputs of fwd are not needed by bwd but the descriptor check wants non-null
        s.logits, s.chars = fws.data_ptr(), t['chars'].data_ptr()
        # the backward loop is graph-cached on its pointers too: stage dlogits in, write every gradient into slot buffers
        slot = lease.slot
        dl = slot.buf('dlogits', tuple(dlogits.shape), torch.float32, K.device)
        dl.copy_(dlogits)
        g = LasSpellerGrads()
        g.dlogits = dl.data_ptr()
        gbufs = [slot.buf('g_' + name, tuple(p.shape), torch.float32, K.device) for name, p in zip(SPELLER_PARAM_ORDER, params)]
        for name, tt in zip(SPELLER_PARAM_ORDER, gbufs):
            setattr(g, 'd_' + name, tt.data_ptr())
        # d_w_ih0 is written in three column/row pieces that together cover it; no zero-init needed
        dKb = slot.buf('dK', tuple(K.shape), torch.float32, K.device)
        dVb = slot.buf('dV', tuple(V.shape), torch.float32, K.device)
        g.dK, g.dV = dKb.data_ptr(), dVb.data_ptr()
        # Parameter gradients beside the encoder's backward (the decoder's share of the "backward overlap" above): when the caller's
        # parameters keep their gradients in the reducer's buckets, only the loop + dK / dV are enqueued here; the batched parameter
        # gradients (a dozen split-K GEMMs, column sums and small FFMA GEMMs, ~0.5 ms at B = 96, L = 300 that nothing downstream
        # waits for) are queued like an encoder layer's weight gradients: the top encoder layer's backward issues them on the second
        # stream beside its BPTT kernel, they accumulate straight into p.grad and report to the reducer.  LAS_BWD_SPELLER_OVERLAP=0: off.
        wrefs = ctx.wrefs
        ovl = None
        if wrefs is not None and use_tc and os.environ.get('LAS_BWD_SPELLER_OVERLAP', '1') != '0' and _overlap_ok(wrefs) \
                and all(w.shape == p.shape for w, p in zip(wrefs, params)):
            ovl = _overlap_state(K.device)
        if ovl is None:
            check(lib.las_speller_bwd_f32(C.byref(s), C.byref(g), stream_ptr()), 'speller_bwd')
            grads = [t_.clone() for t_ in gbufs]           # the caller (autograd) owns what it receives
            dK, dV = dKb.clone(), dVb.clone()
            lease.release()
            return (dK, dV, None, None, None, None, None, None, None, None, None, None, None, None, *grads)
        _overlap_begin(ovl)
        check(lib.las_speller_bwd_phases_f32(C.byref(s), C.byref(g), 1, stream_ptr()), 'speller_bwd (loop)')
        dK, dV = dKb.clone(), dVb.clone()

        def run(_max_ctas, s=s, g=g, keep=keep, gbufs=gbufs, wrefs=wrefs, lease=lease, held=(t, params, dl)):
            # side stream current; ordered behind the loop by the caller (BPTT launch-start hand-off or wait_stream at the end of the pass)
            check(lib.las_speller_bwd_phases_f32(C.byref(s), C.byref(g), 2, stream_ptr()), 'speller_bwd (parameter gradients)')
            for w, gb in zip(wrefs, gbufs):
                w.grad.add_(gb)
            for w in wrefs:                                # stands in for the post-accumulate-grad hook (bucket all-reduce)
                ready = getattr(w, '_las_grad_ready', None)
                if ready is not None:
                    ready(w)
            # the slot's buffers are read on the side stream: the slot goes back to the pool now, but whoever takes it next writes on
            # the main stream, which joins the side stream at the end of this backward pass (_overlap_finish) before anything else runs
            _overlap_state(gbufs[0].device).keepalive.append((gbufs, held, keep))
            lease.release()

        for w in wrefs:                                    # the reducer's autograd hook must not count the None returned below
            w._las_deferred = True
        ovl.pending.append(run)
        return (dK, dV, None, None, None, None, None, None, None, None, None, None, None, None, *([None] * len(gbufs)))


def speller_loop(K, V, enc_lens, params, *, steps, heads, sos_idx, pad_idx, training, dec_y=None, use_gold=None,
                 drop0=None, drop1=None, init_force=False, defer_param_grads=False):
    """defer_param_grads: the caller guarantees that `params` are used by nothing else in the autograd graph of this step (true for the
    LAS Speller; NOT for the Rewriter, whose char_emb also embeds the encoder's input), so their gradients may be accumulated into
    p.grad on the second stream beside the encoder's backward instead of being returned through autograd (SpellerFunction.backward)."""
    wrefs = tuple(params) if (defer_param_grads and training) else None
    return SpellerFunction.apply(K, V, enc_lens, dec_y, use_gold, drop0, drop1, int(steps), int(heads), int(sos_idx),
                                 int(pad_idx), bool(training), bool(init_force), wrefs, *params)
