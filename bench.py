#!/usr/bin/env python
"""bench.py -- LAS teacher-forced training throughput (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus 1 --steps K --warmup W                 # our arm (CUDA path through the public nn.Module API)
    python bench.py --impl reference --gpus 1 --steps K --warmup W  # reference arm: the unmodified reference on the host cores
    torchrun ... bench.py --gpus N ...                            # N > 1: one rank per GPU, NCCL gradient all-reduce

A "step" = one pass of the hot path over one synthetic batch: Listener + Speller forward under teacher forcing, masked CE,
backward (BPTT), unscale + clip + AdamW(amsgrad) update [+ gradient all-reduce].  Workload = BASELINE.json configs[1]:
best base-LAS (hid 512, 1 LSTM + 3 pLSTM bidirectional, att_proj 256, dec 512/256), batch 96 per GPU, T=1600, L=300.
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, 'attention-based-e2e-asr-dnn_b200')
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np   # noqa: E402
import torch         # noqa: E402

METRIC = 'las_train_utterances_per_sec'
UNIT = 'utterances/s'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=96, help='utterances per GPU')
    ap.add_argument('--T', type=int, default=1600)
    ap.add_argument('--L', type=int, default=300)
    ap.add_argument('--config', default='best')
    ap.add_argument('--cpu-sample-batch', type=int, default=8, help='utterances of the CPU-baseline sample (about 10 s of CPU work per step at 8)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-dropout', action='store_true', help='run the model with every dropout at 0 (default: the yml values 0.3/0.3/0.35/0.3)')
    ap.add_argument('--no-gpu-reference', action='store_true', help='skip timing the unmodified reference on the GPU (gpu_reference)')
    ap.add_argument('--precision', default='bf16', choices=['bf16', 'fp32'],
                    help="bf16: the AMP path (gate GEMMs on the tensor pipe, like the reference's autocast runs); fp32: parity mode")
    ap.add_argument('--no-greedy', action='store_true', help='skip the greedy-decode leg (BASELINE configs[3]) reported under "greedy"')
    ap.add_argument('--no-fp32-leg', action='store_true', help='skip the two fp32-parity-mode steps reported under "fp32_mode"')
    ap.add_argument('--no-rewriter', action='store_true', help='skip the Rewriter leg (BASELINE configs[4]) reported under "rewriter"')
    ap.add_argument('--rewriter-batch', type=int, default=64)
    ap.add_argument('--greedy-batch', type=int, default=256)
    ap.add_argument('--greedy-T', type=int, default=3000)
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm=d.get('hbm_gbs', 6650.0), tf_burst=d.get('bf16_tflops', 1590.0),
                    tf_sustained=d.get('bf16_tflops_sustained', 1400.0), source='measured')
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source='fallback')


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-i', str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(',')])
            except Exception:
                pass
            self._stop_evt.wait(0.05)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=sorted(reasons),
                    samples=len(sm))


def ncu_dram_bytes(kernel_prefix):
    """dram read + write bytes per launch of the first row whose kernel name starts with `kernel_prefix` in the newest committed
    `ncu --set full` extract under profiles/ (columns: kernel, dram_rd [GB], dram_wr [GB], ...), or None."""
    for name in ('ncu_full_r2_final_kernels.csv', 'ncu_full_r2_kernels.csv', 'ncu_full_r1_final_kernels.csv', 'ncu_full_r1_kernels.csv'):
        path = os.path.join(ROOT, 'profiles', name)
        if not os.path.exists(path):
            continue
        for line in open(path):
            if line.startswith('#') or line.startswith('kernel,'):
                continue
            if kernel_prefix in line and '>,' in line:
                rest = line.split('>,', 1)[1].split(',')          # columns after the kernel name: dram_rd [GB], dram_wr [GB], ...
                try:
                    return dict(bytes=(float(rest[0]) + float(rest[1])) * 1e9, source='profiles/' + name)
                except (ValueError, IndexError):
                    continue
    return None


def oracle_cpu_step(cfg_name, B, T, L, seed=11785):
    """One teacher-forced train step (fwd + bwd + AdamW-amsgrad) of the reference algorithm's CPU port (oracle/): the checker,
    timed here as the CPU baseline.  Returns seconds."""
    from oracle import golden_util as gu
    from oracle import las_oracle as orc
    torch.set_num_threads(os.cpu_count())
    cfg = gu.get_config(cfg_name)
    sd = gu.make_state_dict(cfg, seed)
    x, lx, y = gu.make_inputs(seed + 1, B, T, L)
    p = {k: torch.from_numpy(v.copy()).requires_grad_(True) for k, v in sd.items() if k != 'spell.cls.weight'}
    p['spell.cls.weight'] = p['spell.char_emb.weight']
    lc = cfg['listener_configs']
    t0 = time.perf_counter()
    logits, _ = orc.las_forward(p, torch.from_numpy(x), lx.tolist(), lstm_layers=lc['lstm_layers'], plstm_layers=lc['plstm_layers'],
                                heads=1, training=True, steps=L, dec_y=torch.from_numpy(y), coins=[True] * L)
    loss = orc.masked_ce_loss(logits, torch.from_numpy(y), [L] * B)
    loss.backward()
    names = [k for k in p if k != 'spell.cls.weight']
    params = [p[k].detach() for k in names]
    grads = [p[k].grad for k in names]
    orc.optimizer_step(params, grads, [dict() for _ in names], lr=5e-4, weight_decay=5e-6)
    return time.perf_counter() - t0


def attn_step_replay_us(lib, B, T, P, dev, nrep=50):
    """Average duration of the fused attention-step kernels (fwd, bwd) at one shape, issued the way the decoder loop issues
    them: `nrep` launches captured in a CUDA graph and replayed back to back (K/V stay L2-warm at the train shape, as in
    the loop), timed with CUDA events on the replay stream.  Returns (us_fwd, us_bwd, algorithmic bytes per launch)."""
    import ctypes as C
    from las_b200 import _lib
    from las_b200._lib import LasAttnStep
    q = torch.randn(B, P, device=dev); K = torch.randn(B, T, P, device=dev); V = torch.randn(B, T, P, device=dev)
    lens = torch.full((B,), T, dtype=torch.int32, device=dev)
    ctx = torch.empty(B, P, device=dev); w = torch.empty(B, 1, T, device=dev)
    dctx = torch.randn(B, P, device=dev); dq = torch.empty(B, P, device=dev); de = torch.empty(B, 1, T, device=dev)
    d = LasAttnStep()
    d.q, d.ld_q = q.data_ptr(), P
    d.K, d.V, d.lens = K.data_ptr(), V.data_ptr(), lens.data_ptr()
    d.w, d.ld_w = w.data_ptr(), T
    d.ctx, d.ld_ctx = ctx.data_ptr(), P
    d.dctx, d.ld_dctx = dctx.data_ptr(), P
    d.dq, d.ld_dq, d.dq_accumulate = dq.data_ptr(), P, 0
    d.de = de.data_ptr()
    d.B, d.T, d.P, d.heads = B, T, P, 1
    d.scale = float(P ** 0.5)
    out = []
    for name, fn in (('fwd', lib.las_attn_step_fwd_f32), ('bwd', lib.las_attn_step_bwd_f32)):
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            _lib.check(fn(C.byref(d), side.cuda_stream), name)
            torch.cuda.synchronize()
            # thread_local: the NCCL watchdog thread keeps polling events while this thread captures
            with torch.cuda.graph(g, stream=side, capture_error_mode='thread_local'):
                for _ in range(nrep):
                    _lib.check(fn(C.byref(d), side.cuda_stream), name)
        torch.cuda.synchronize()
        ts = []
        for i in range(8):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record()
            torch.cuda.synchronize()
            if i >= 3:
                ts.append(1e3 * e0.elapsed_time(e1) / nrep)
        out.append(float(np.median(ts)))
        del g
    return out[0], out[1], 2.0 * B * T * P * 4


def ref_runner(cmd_args, timeout=900):
    """oracle/ref_runner.py in its own process (the reference's top-level package is called `src`, like the drop-in shim, so it
    never shares a process with the product package).  Returns the JSON dict it prints, or None."""
    staged = os.path.isdir(os.path.join(ROOT, 'oracle', '_ref', 'src')) or os.path.isdir('/root/reference/src')
    if not staged:
        return None
    env = {k: v for k, v in os.environ.items() if k not in ('RANK', 'LOCAL_RANK', 'WORLD_SIZE', 'MASTER_ADDR', 'MASTER_PORT', 'PYTHONPATH')}
    try:
        out = subprocess.run([sys.executable, os.path.join(ROOT, 'oracle', 'ref_runner.py')] + [str(a) for a in cmd_args], env=env,
                             capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    except subprocess.TimeoutExpired:
        return None
    for line in reversed(out.stdout.splitlines()):
        if line.startswith('{'):
            try:
                return json.loads(line)
            except ValueError:
                continue
    sys.stderr.write('ref_runner %s failed:\n%s\n' % (cmd_args, out.stderr[-1500:]))
    return None


def reference_cpu_step(cfg_name, B, T, L, steps=1, warmup=0):
    """Seconds per train step of the UNMODIFIED reference on the host cores (all threads), or None when it is not staged."""
    d = ref_runner(['bench', '--device', 'cpu', '--config', cfg_name, '--B', B, '--T', T, '--L', L, '--steps', steps, '--warmup', warmup])
    return None if d is None else d['ms_per_step'] / 1e3


def run_reference(args):
    """Reference arm: the UNMODIFIED reference (pure Python on PyTorch; staged by oracle/build_ref.sh into the git-ignored oracle/_ref so
    that it travels to the GPU box) on all host cores, through its own nn.Module API and the trainer's step sequence
    (oracle/ref_runner.py bench), on a bounded sample of the workload.  Falls back to the CPU port (oracle/las_oracle.py, pinned to the
    reference by tests/golden) only when the staged copy is absent."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    per_utt = reference_cpu_step(args.config, 1, args.T, args.L)
    if per_utt is not None:
        # size the per-step sample so that the K timed steps take about two and a half minutes at most
        B = int(max(1, min(args.cpu_sample_batch, 150.0 / (max(args.steps, 1) * max(per_utt, 1e-3)))))
        d = ref_runner(['bench', '--device', 'cpu', '--config', args.config, '--B', B, '--T', args.T, '--L', args.L, '--steps', args.steps,
                        '--warmup', min(args.warmup, 1)], timeout=1500)
        if d is not None:
            ms, val = d['ms_per_step'], d['utt_per_s']
            sample = f'B={B} utterances of the same T={args.T}, L={args.L} workload per step (yml dropouts on)'
            out = dict(impl='reference', metric=METRIC, value=val, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                       ms_per_step=ms, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32', data='synthetic',
                       config=dict(workload=f'{args.config} base-LAS teacher-forced train step, T={args.T}, L={args.L}', sample=sample),
                       cpu_baseline=dict(value=val, unit=UNIT, cores=d['threads'], kind='reference', sample=sample),
                       e2e=dict(value=val, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
            print(json.dumps(out), flush=True)
            return
    for i in range(max(1, min(args.warmup, 1))):
        oracle_cpu_step(args.config, 1, max(args.T // 8, 8), max(args.L // 8, 2))      # warm the allocator / thread pool
    # size the per-step sample so that the K timed steps take about two and a half minutes at most: probe one utterance first
    per_utt = oracle_cpu_step(args.config, 1, args.T, args.L)
    B = int(max(1, min(args.cpu_sample_batch, 150.0 / (max(args.steps, 1) * max(per_utt, 1e-3)))))
    times = []
    for i in range(args.steps):
        times.append(oracle_cpu_step(args.config, B, args.T, args.L))
    ms = 1e3 * float(np.mean(times))
    val = B / (ms / 1e3)
    sample = f'B={B} utterances of the same T={args.T}, L={args.L} workload per step'
    out = dict(impl='reference', metric=METRIC, value=val, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
               ms_per_step=ms, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32', data='synthetic',
               config=dict(workload=f'{args.config} base-LAS teacher-forced train step, T={args.T}, L={args.L}', sample=sample),
               cpu_baseline=dict(value=val, unit=UNIT, cores=os.cpu_count(), kind='port', sample=sample),
               e2e=dict(value=val, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(out), flush=True)


def main():
    args = parse()
    if args.impl == 'reference':
        return run_reference(args)

    import torch.distributed as dist
    import las_b200
    from las_b200 import _lib
    from las_b200.ddp import BucketedGradReducer
    from las_b200.models import ListenAttendSpell
    from las_b200.optim import FusedAdamW
    from las_b200 import configs as gu        # config table + seeded synthetic inputs (the product arm never touches oracle/)

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py (impl ours) needs a GPU: las_b200 has no CPU fallback')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    lib = _lib.load()
    _lib.check(lib.las_init(local), 'las_init')

    B, T, L = args.batch, args.T, args.L
    # the config the bench names runs with ITS dropouts (config/sample-attention.yml:50-65: 0.3 / 0.3 / 0.35, decoder 0.3): mask
    # draws, masked-output writes and the decoder's per-step mask staging are inside the timed region
    drop = {} if args.no_dropout else dict(init_dropout=0.3, mid_dropout=0.3, final_dropout=0.35, dec_lstm_dropout=0.3)
    cfg = gu.get_config(args.config, **drop)
    torch.manual_seed(11785)                     # the reference's seed (config/sample-attention.yml:11), same on every rank
    model = ListenAttendSpell(**cfg).to(dev).train()
    if world > 1:
        for p in model.parameters():
            dist.broadcast(p.data, 0)
    opt = FusedAdamW(model.parameters(), lr=5e-4, weight_decay=5e-6, amsgrad=True)
    reducer = BucketedGradReducer(list(model.named_parameters()), world_size=world)
    x_np, lx_np, y_np = gu.make_inputs(11785 + rank, B, T, L)
    x_host = torch.from_numpy(x_np).pin_memory()
    y_host = torch.from_numpy(y_np).pin_memory()
    lx = torch.from_numpy(lx_np)                 # CPU int64, never moved (src/train.py:127)
    x_dev, y_dev = x_host.to(dev), y_host.to(dev)
    from las_b200.loss import masked_ce
    ly_cpu = torch.full((B,), L, dtype=torch.int64)      # every target position is non-pad in the synthetic batch
    V = cfg['speller_configs']['dec_vocab_size']
    scale = 65536.0                              # GradScaler's initial scale (torch amp/grad_scaler.py)

    def step(x, y):
        reducer.zero_grad()
        # like src/train.py:130-137: forward + loss under autocast (which selects our tensor-pipe kernels)
        with torch.autocast('cuda', dtype=torch.bfloat16, enabled=(args.precision == 'bf16')):
            logits, _att = model(x, lx, y, 1.0, False)                   # tf_rate 1.0 (README stage 1)
        loss, _ppl = masked_ce(logits, y, ly_cpu)                         # the trainer's masked CE (src/train.py:117-136), fused
        (loss * scale).backward()
        reducer.finish()
        opt.step_fused(inv_scale=1.0 / (scale * world), max_norm=5.0)
        return loss

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        sync()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    min_warm = int(os.environ.get('LAS_BENCH_MIN_WARMUP', '3'))      # timing rule: W >= 3 (lowered only for ncu captures)
    for _ in range(max(args.warmup, min_warm)):
        step(x_dev, y_dev)
    sync()

    # ---- device-resident throughput (inputs already in HBM) ----
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    PROF_KINDS = dict(gemm_gates=0, gemm_other=1, rec_fwd=2, rec_bwd=3, attn_fwd=4, attn_bwd=5, adam=6, speller_fwd=7, speller_bwd=8,
                      gemm_gates_side=9)
    lib.las_prof_reset()
    # top-level kinds only (gate GEMMs, recurrence, optimizer, whole decoder loop): profiling the kernels INSIDE the decoder loop
    # would disable its CUDA-graph replay; the attention step is timed in a separate, untimed-for-throughput step below
    lib.las_prof_enable(0b1111001101 if rank == 0 else 0)
    las_b200.reset_launch_count()
    ms_total = timed(lambda: step(x_dev, y_dev), args.steps)
    launches = las_b200.launch_count()
    lib.las_prof_enable(0)
    clocks = sampler.stop() if sampler else None
    ms_step = ms_total / args.steps
    value = world * B / (ms_step / 1e3)

    import ctypes as C
    prof = {}
    if rank == 0:
        for name, kind in PROF_KINDS.items():
            ms, n, work = C.c_double(), C.c_longlong(), C.c_double()
            _lib.check(lib.las_prof_collect(kind, C.byref(ms), C.byref(n), C.byref(work)), 'prof_collect')
            prof[name] = dict(ms_per_step=ms.value / args.steps, launches_per_step=n.value / args.steps, work_per_step=work.value / args.steps)
        lib.las_prof_reset()
    # one extra (untimed) step with the decoder's inner kernels profiled on rank 0: attention step fwd / bwd, small GEMMs.
    # EVERY rank runs the step (it contains the gradient all-reduce); only rank 0 records.
    lib.las_prof_enable(((1 << 1) | (1 << 4) | (1 << 5)) if rank == 0 else 0)
    step(x_dev, y_dev)
    torch.cuda.synchronize()
    lib.las_prof_enable(0)
    if rank == 0:
        for name in ('gemm_other', 'attn_fwd', 'attn_bwd'):
            ms, n, work = C.c_double(), C.c_longlong(), C.c_double()
            _lib.check(lib.las_prof_collect(PROF_KINDS[name], C.byref(ms), C.byref(n), C.byref(work)), 'prof_collect')
            prof[name] = dict(ms_per_step=ms.value, launches_per_step=float(n.value), work_per_step=work.value)
        lib.las_prof_reset()

    # gate-GEMM roofline: in the timed region most of these GEMMs run as 128-row time tiles BESIDE the recurrence kernels,
    # where a launch's duration says nothing about the kernel.  Two extra (untimed) steps with that pipelining switched off run the
    # same kernel as full-size launches that own the GPU; EVERY rank runs them (gradient all-reduce inside), rank 0 records.
    saved_env = {k: os.environ.get(k) for k in ('LAS_FWD_PIPELINE', 'LAS_BWD_PIPELINE')}
    os.environ['LAS_FWD_PIPELINE'] = '0'
    os.environ['LAS_BWD_PIPELINE'] = '0'
    step(x_dev, y_dev)
    torch.cuda.synchronize()
    lib.las_prof_reset()
    lib.las_prof_enable(((1 << 0) | (1 << 9)) if rank == 0 else 0)
    NUNP = 2
    for _ in range(NUNP):
        step(x_dev, y_dev)
    torch.cuda.synchronize()
    lib.las_prof_enable(0)
    for k, v in saved_env.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v
    if rank == 0:
        for name in ('gemm_gates', 'gemm_gates_side'):
            ms, n, work = C.c_double(), C.c_longlong(), C.c_double()
            _lib.check(lib.las_prof_collect(PROF_KINDS[name], C.byref(ms), C.byref(n), C.byref(work)), 'prof_collect')
            prof[name + '_unpipelined'] = dict(ms_per_step=ms.value / NUNP, launches_per_step=n.value / NUNP, work_per_step=work.value / NUNP)
        lib.las_prof_reset()

    attn_replay = None
    if rank == 0:
        Pq = cfg['speller_configs']['att_proj_dim']
        uf, ub, byts = attn_step_replay_us(lib, B, T // 8, Pq, dev)
        attn_replay = dict(us_fwd=uf, us_bwd=ub, bytes_per_launch=byts)

    # ---- end to end: host (pinned) inputs -> device every step, loss read back every step ----
    def e2e_step():
        xd = x_host.to(dev, non_blocking=True)
        yd = y_host.to(dev, non_blocking=True)
        return float(step(xd, yd).item())
    e2e_step()
    ms_e2e = timed(e2e_step, args.steps) / args.steps
    e2e_value = world * B / (ms_e2e / 1e3)
    h2d = x_host.numel() * x_host.element_size() + y_host.numel() * y_host.element_size()

    # ---- the same train step in fp32 parity mode (the mode that carries the 1e-4 evidence): FFMA GEMMs + fp32 recurrence ----
    fp32_mode = None
    if rank == 0 or world > 1:
        from las_b200.precision import set_precision
        if args.precision == 'bf16' and not args.no_fp32_leg:
            def step32():
                reducer.zero_grad()
                set_precision('fp32')
                try:
                    logits, _att = model(x_dev, lx, y_dev, 1.0, False)
                    loss, _ppl = masked_ce(logits, y_dev, ly_cpu)
                    (loss * scale).backward()
                finally:
                    set_precision('auto')
                reducer.finish()
                opt.step_fused(inv_scale=1.0 / (scale * world), max_norm=5.0)
            step32()
            ms32 = timed(step32, 2) / 2
            fp32_mode = dict(ms_per_step=ms32, value=world * B / (ms32 / 1e3), unit=UNIT, note='fp32 parity mode (LAS_PRECISION=fp32), 2 timed steps')

    # ---- Rewriter (BASELINE configs[4], src/lmtrain.py:95-253, config/rewriter.yml): char-to-char LM, tf_rate 0.5 ----
    rewriter = None
    if not args.no_rewriter:
        from las_b200.lm import Rewriter
        rcfg = gu.get_rewriter_config('rw_yml')
        Br, Tr = args.rewriter_batch, 200
        torch.manual_seed(11785)
        rw = Rewriter(**rcfg).to(dev).train()
        if world > 1:
            for p in rw.parameters():
                dist.broadcast(p.data, 0)
        rw_red = BucketedGradReducer(list(rw.named_parameters()), world_size=world, bucket_key=lambda n: 'rewriter')
        rw_opt = FusedAdamW(rw.parameters(), lr=1e-3, weight_decay=5e-6, amsgrad=True)
        xr_np, lxr_np, yr_np = gu.make_token_inputs(777 + rank, Br, Tr, Tr)
        xr, yr, lxr = torch.from_numpy(xr_np).to(dev), torch.from_numpy(yr_np).to(dev), torch.from_numpy(lxr_np)
        lyr = torch.full((Br,), Tr, dtype=torch.int64)

        def rw_step():
            rw_red.zero_grad()
            with torch.autocast('cuda', dtype=torch.bfloat16, enabled=(args.precision == 'bf16')):
                lg, _ = rw(xr, lxr, yr, 0.5)
            loss, _ = masked_ce(lg, yr, lyr)
            (loss * scale).backward()
            rw_red.finish()
            rw_opt.step_fused(inv_scale=1.0 / (scale * world), max_norm=5.0)
        for _ in range(8):                  # the decoder-loop graphs (forward / backward, three segments each) settle within a few steps
            rw_step()
        ms_rw = timed(rw_step, 10) / 10
        rewriter = dict(metric='rewriter_train_sequences_per_sec', value=world * Br / (ms_rw / 1e3), unit='sequences/s', ms_per_step=ms_rw,
                        config=dict(workload=f'Rewriter (config/rewriter.yml dims: emb 256, 2x BiLSTM 256, P 128 x 4 heads, dec 256/128, yml dropouts), '
                                             f'batch {Br}/GPU, Tx = L = {Tr}, tf_rate 0.5 (never takes effect: src/lmtrain.py:231), fwd+bwd+AdamW'))
        del rw, rw_red, rw_opt
        torch.cuda.empty_cache()

    # ---- greedy decoding (BASELINE configs[3]): eval mode, CHR_MAX_STEPS = 600 steps always (src/models.py:315), no collective ----
    greedy = None
    if not args.no_greedy:
        del x_dev, y_dev
        reducer.zero_grad()
        torch.cuda.empty_cache()
        Bg, Tg = args.greedy_batch, args.greedy_T
        xg_np, lxg_np, _ = gu.make_inputs(4242 + rank, Bg, Tg, 4)
        xg_host = torch.from_numpy(xg_np).pin_memory()
        xg = xg_host.to(dev)
        lxg = torch.from_numpy(lxg_np)
        model.eval()
        steps_g = cfg['speller_configs']['CHR_MAX_STEPS']

        def decode(xin):
            with torch.no_grad(), torch.autocast('cuda', dtype=torch.bfloat16, enabled=(args.precision == 'bf16')):
                lg, _ = model(xin, lxg)
            return lg
        for _ in range(2):
            decode(xg)
        lib.las_prof_reset()
        lib.las_prof_enable((1 << 4) if rank == 0 else 0)
        n_g = 3
        ms_g = timed(lambda: decode(xg), n_g) / n_g
        lib.las_prof_enable(0)
        ga = dict(ms=0.0, n=0, work=0.0)
        if rank == 0:
            import ctypes as C2
            ms_, n_, w_ = C2.c_double(), C2.c_longlong(), C2.c_double()
            _lib.check(lib.las_prof_collect(4, C2.byref(ms_), C2.byref(n_), C2.byref(w_)), 'prof_collect')
            ga = dict(ms=ms_.value, n=n_.value, work=w_.value)
            lib.las_prof_reset()

        def decode_e2e():
            lg = decode(xg_host.to(dev, non_blocking=True))
            return lg.argmax(-1).to(torch.int16).cpu()        # transcripts back on the host
        decode_e2e()
        ms_ge = timed(decode_e2e, 2) / 2
        ga_replay = None
        if rank == 0:
            uf, _ub, byts = attn_step_replay_us(lib, Bg, Tg // 8, cfg['speller_configs']['att_proj_dim'], dev)
            ga_replay = dict(us_per_launch=uf, bytes_per_launch=byts, gbs=byts / uf / 1e3)
        greedy = dict(metric='las_greedy_decode_chars_per_sec', value=world * Bg * steps_g / (ms_g / 1e3), unit='chars/s',
                      ms_per_batch=ms_g, e2e_value=world * Bg * steps_g / (ms_ge / 1e3),
                      config=dict(workload=f'{args.config} base-LAS greedy decode, batch {Bg}/GPU, T={Tg}, {steps_g} steps (CHR_MAX_STEPS), eval mode'),
                      attn_step=(dict(us_per_launch=1e3 * ga['ms'] / ga['n'], bytes_per_launch=ga['work'] / ga['n'],
                                      gbs=(ga['work'] / 1e9) / (ga['ms'] / 1e3) if ga['ms'] > 0 else 0.0,
                                      method='CUDA events around every launch inside the decode loop (includes the per-launch event/launch gap)')
                                 if ga['n'] > 0 else
                                 dict(us_per_launch=None, launches=0,
                                      note='no separate attention launches: the attention phase runs inside the persistent decoder kernel '
                                           '(csrc/decoder_persist.cu); see attn_step_replay for the stand-alone kernel at this shape')),
                      attn_step_replay=ga_replay)
        model.train()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    gemm_traffic = ncu_dram_bytes('gemm_bf16_tc_kernel<0, 0, ')
    attn_traffic = ncu_dram_bytes('attn_step_split_kernel<0')
    gg = prof['gemm_gates_unpipelined']
    gg_timed = prof['gemm_gates']
    tf_achieved = (gg['work_per_step'] / 1e12) / (gg['ms_per_step'] / 1e3) if gg['ms_per_step'] > 0 else 0.0
    gs = prof.get('gemm_gates_side', dict(ms_per_step=0.0, work_per_step=0.0))
    # the weight-gradient GEMMs of layers 1-3 run on a second stream BESIDE the BPTT kernel of the layer below, on the 52 SMs it leaves
    # free (functional.py, backward overlap): their time is hidden, not comparable with a whole-GPU peak, and reported apart
    roofline = dict(kernel='lstm input-gate GEMMs that own the GPU (fwd + dgrad all layers, wgrad of the base layer)' if gs['ms_per_step'] > 0
                    else 'lstm input-gate GEMMs (fwd + dgrad + wgrad, all layers)', bound='tensor', achieved=tf_achieved,
                    peak=pk['tf_sustained'], unit='TFLOP/s', frac=tf_achieved / pk['tf_sustained'],
                    traffic=(gemm_traffic or {}).get('bytes'),
                    traffic_note='dram read+write of the largest launch (layer-1 forward, M=76800 N=4096 K=2048: 1.29 TFLOP, 1.59 GB algorithmic), '
                                 'read at run time from the committed ncu --set full extract ' + str((gemm_traffic or {}).get('source')),
                    peak_source=pk['source'] + ' bf16 sustained', ms_per_step=gg['ms_per_step'], flops_per_step=gg['work_per_step'],
                    measured='CUDA events around every launch of the kernel in 2 extra steps after the timed region, with LAS_FWD_PIPELINE=0 and '
                             'LAS_BWD_PIPELINE=0: full-size launches that own the GPU.  In the timed region the same kernel ran as in '
                             '`in_timed_region`: part of it as 128-row time tiles beside the recurrence kernels (`beside_recurrence`)',
                    in_timed_region=dict(owning_gpu=dict(ms_per_step=gg_timed['ms_per_step'], flops_per_step=gg_timed['work_per_step'],
                                                         launches_per_step=gg_timed['launches_per_step'])),
                    beside_recurrence=dict(ms_per_step=gs['ms_per_step'], flops_per_step=gs['work_per_step'], max_ctas=52,
                                           note="wgrad GEMMs of layers 1-3 (capped to the SMs the BPTT kernel leaves free) and the time tiles of the next layer's gate GEMM / this layer's dX GEMM issued behind the recurrence kernels' progress counters: overlapped with those kernels"))
    af = prof['attn_fwd']
    # attention step: the kernel's own duration = graph-replayed back-to-back launches at the workload shape (the decoder loop
    # replays it the same way); the per-launch-evented in-loop figure is kept beside it (it includes launch/event gaps)
    at_gbs = attn_replay['bytes_per_launch'] / attn_replay['us_fwd'] / 1e3
    at_gbs_bwd = attn_replay['bytes_per_launch'] / attn_replay['us_bwd'] / 1e3
    attn_roofline = dict(kernel='fused attention step fwd (energy+masked softmax+context), single-pass T-split', bound='hbm',
                         achieved=at_gbs, peak=pk['hbm'], unit='GB/s', frac=at_gbs / pk['hbm'], frac_of_8tbs=at_gbs / 8000.0,
                         traffic=(attn_traffic or {}).get('bytes'),
                         traffic_note='dram read+write per launch with a flushed L2 (ncu --set full extract ' + str((attn_traffic or {}).get('source')) +
                                      '): K and V are read exactly once; in the decoder loop they are L2 hits',
                         peak_source=pk['source'],
                         us_per_launch=attn_replay['us_fwd'], bytes_per_launch=attn_replay['bytes_per_launch'],
                         bwd=dict(achieved=at_gbs_bwd, frac=at_gbs_bwd / pk['hbm'], frac_of_8tbs=at_gbs_bwd / 8000.0, us_per_launch=attn_replay['us_bwd']),
                         method='50 launches captured in a CUDA graph, replayed, CUDA events on the replay stream; K/V (39 MB) L2-warm as in the loop',
                         in_loop_evented_us_per_launch=(1e3 * af['ms_per_step'] / af['launches_per_step'] if af['launches_per_step'] > 0 else None),
                         in_loop_note=(None if af['launches_per_step'] > 0 else 'the forward decoder loop launches no attention kernel in AMP mode: its '
                                       'attention phase is part of dec_persist_fwd_kernel (decoder_step_us.fwd); this kernel serves the backward loop, '
                                       'fp32 mode and the stand-alone module'),
                         note='K/V of one batch fit in the 126 MB L2, so algorithmic GB/s can exceed what DRAM alone would give')
    T_total = prof['rec_fwd']['work_per_step']
    rec = dict(fwd_us_per_timestep=1e3 * prof['rec_fwd']['ms_per_step'] / max(T_total, 1),
               bwd_us_per_timestep=1e3 * prof['rec_bwd']['ms_per_step'] / max(T_total, 1), timesteps_per_step=T_total)

    cpu = None
    gpu_ref = None
    if world == 1 and not args.no_cpu_baseline:
        Bs = args.cpu_sample_batch
        sec = reference_cpu_step(args.config, Bs, T, L)
        if sec is not None:
            cpu = dict(value=Bs / sec, unit=UNIT, cores=os.cpu_count(), kind='reference',
                       sample=f'one train step (fwd+bwd+clip+AdamW, yml dropouts) of the unmodified reference (oracle/_ref) at B={Bs}, T={T}, L={L}: {sec:.1f} s')
        else:
            sec = oracle_cpu_step(args.config, Bs, T, L)
            cpu = dict(value=Bs / sec, unit=UNIT, cores=os.cpu_count(), kind='port',
                       sample=f'one fwd+bwd+AdamW step of the CPU port (oracle/) at B={Bs}, T={T}, L={L}: {sec:.1f} s')
    if world == 1 and not args.no_gpu_reference:
        # secondary comparator (SURVEY 2.1 / BASELINE.md section 3): the unmodified reference on this B200 through torch's cuDNN / cuBLAS
        # path, same batch and lengths -- fp32, and under autocast as its trainer runs it (src/train.py:130, fp16 by default)
        torch.cuda.empty_cache()
        gpu_ref = {}
        for amp in ('none', 'bf16', 'fp16'):
            d = ref_runner(['bench', '--device', 'cuda', '--amp', amp, '--config', args.config, '--B', B, '--T', T, '--L', L, '--steps', 3,
                            '--warmup', 2], timeout=600)
            if d is not None:
                gpu_ref['fp32' if amp == 'none' else amp + '_autocast'] = dict(ms_per_step=d['ms_per_step'], value=d['utt_per_s'], unit=UNIT)
        if not gpu_ref:
            gpu_ref = None
        else:
            gpu_ref['note'] = ('unmodified reference (oracle/_ref) on the same GPU: nn.LSTM -> cuDNN, per-step Python decoder loop with a '
                               'blocking .cpu() per step (src/models.py:377), GradScaler + clip_grad_norm_ + torch AdamW; 3 timed steps')

    out = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=max(args.warmup, min_warm), ms_per_step=ms_step,
               higher_is_better=True, scaling='weak', vs_baseline=None, dtype=('bf16' if args.precision == 'bf16' else 'f32'), data='synthetic',
               config=dict(workload=f'{args.config} base-LAS teacher-forced train step (fwd+bwd+unscale/clip/AdamW-amsgrad), '
                                    f'batch {B}/GPU, T={T}, L={L}, tf_rate=1.0, dropouts ' + ('off' if args.no_dropout else '0.3/0.3/0.35 + decoder 0.3 (yml)'),
                           global_batch=B * world, parallelism=f'dp{world}',
                           l2_policy='inputs+activations per step (>2.5 GB) exceed the 126 MB L2; no explicit flush'),
               clocks=clocks, e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=4, ms_per_step=ms_e2e),
               gpu_launches=int(launches), roofline=roofline, attn_roofline=attn_roofline, recurrence=rec,
               kernel_ms_per_step={k: round(v['ms_per_step'], 3) for k, v in prof.items()}, greedy=greedy, cpu_baseline=cpu,
               gpu_reference=gpu_ref, fp32_mode=fp32_mode, rewriter=rewriter,
               decoder_step_us=dict(fwd=1e3 * prof['speller_fwd']['ms_per_step'] / max(L, 1), bwd=1e3 * prof['speller_bwd']['ms_per_step'] / max(L, 1),
                                    note='whole decoder loop / L steps; forward = one persistent cooperative kernel (csrc/decoder_persist.cu)'))
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
