"""
ORACLE -- TEST INFRASTRUCTURE ONLY.  Always run as its own process:  python oracle/ref_runner.py <command> [options]

Drives the UNMODIFIED reference (oracle/_ref staged by oracle/build_ref.sh, or /root/reference) and prints one JSON line:

  bench    one teacher-forced train step the way the reference's trainer runs it (src/train.py:105-190: forward [+ autocast],
           masked CE, [GradScaler] backward, unscale_, clip_grad_norm_(5.0), AdamW(amsgrad) step, zero_grad) on synthetic data of
           BASELINE.json's shape, on the CPU (bench.py --impl reference, cpu_baseline.kind "reference") or on the GPU through
           torch's cuDNN / cuBLAS path (the secondary comparator `gpu_reference`)
  greedy   one greedy decode (eval mode, CHR_MAX_STEPS steps)
  fixture  the reference on a committed fixture (tests/golden/*.npz) with the recorded masks replayed, on any device and under
           autocast: what the AMP parity tests compare the CUDA path's bf16 mode with
  trainer  two batches of the reference's own Trainer.train_epoch + one evaluate_epoch + infer_one_checkpoint on synthetic
           loaders.  With --shim the product package is put on the path first, so `src.models` / `src.modules` resolve to the
           B200-native modules while `src.train` / `src.infer` / `src.utils` stay the reference's: the drop-in run.
"""
from __future__ import annotations

import argparse
import contextlib
import json
import os
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
PKG = os.path.join(ROOT, 'attention-based-e2e-asr-dnn_b200')


def _setup_path(shim: bool):
    sys.path[:] = [p for p in sys.path if os.path.abspath(p or '.') not in (PKG,)]
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from oracle import ref_harness as rh
    ref = rh.reference_root()
    if shim:
        # the INTEGRATION.md recipe: product package first, every other src.* module from the reference checkout
        os.environ['LAS_REFERENCE_SRC'] = os.path.join(ref, 'src')
        for p in (rh.STUBS, PKG):
            sys.path.insert(0, p)
    else:
        for p in (rh.STUBS, ref):
            sys.path.insert(0, p)
    return rh


def _autocast(device: str, amp: str):
    import torch
    if amp == 'none' or device == 'cpu':
        return contextlib.nullcontext()
    return torch.autocast('cuda', dtype=torch.bfloat16 if amp == 'bf16' else torch.float16)


def _sync(device):
    import torch
    if device != 'cpu':
        torch.cuda.synchronize()


def _model(rh, cfg, device, train=True, sd=None):
    import copy
    import torch
    import src.models as models
    m = models.ListenAttendSpell(**copy.deepcopy(cfg))
    if sd is not None:
        m.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in sd.items()})
    m = m.to(device)
    return m.train() if train else m.eval()


def cmd_bench(args, rh):
    import numpy as np
    import torch
    from oracle import golden_util as gu
    over = {}
    if args.dropout:
        over = dict(init_dropout=0.3, mid_dropout=0.3, final_dropout=0.35, dec_lstm_dropout=0.3)     # config/sample-attention.yml:50-65
    cfg = gu.get_config(args.config, **over)
    dev = args.device
    if dev == 'cpu':
        torch.set_num_threads(args.threads or os.cpu_count())
    else:
        torch.backends.cudnn.allow_tf32 = False          # fp32 runs are fp32
        torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(11785)
    model = _model(rh, cfg, dev, train=True)
    opt = torch.optim.AdamW(model.parameters(), lr=5e-4, weight_decay=5e-6, amsgrad=True)          # src/train.py:71-77 + the yml
    use_scaler = args.amp != 'none' and dev != 'cpu'
    scaler = torch.amp.GradScaler('cuda', enabled=use_scaler)
    crit = torch.nn.CrossEntropyLoss(reduction='none')
    x, lx, y = gu.make_inputs(11785, args.B, args.T, args.L)
    x, y, lx = torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev), torch.from_numpy(lx)
    V = cfg['speller_configs']['dec_vocab_size']
    y_mask = torch.ones(args.B * args.L, dtype=torch.int, device=dev)

    def step():
        with _autocast(dev, args.amp):
            logits, _ = model(x, lx, y, 1.0, False)
            loss = (crit(logits.view(-1, V), y.view(-1)) * y_mask).sum() / y_mask.sum()
        scaler.scale(loss).backward()
        scaler.unscale_(opt)
        torch.nn.utils.clip_grad_norm_(model.parameters(), 5.0)
        scaler.step(opt)
        scaler.update()
        opt.zero_grad()
        return float(loss.item())                         # the trainer reads loss.item() every batch (src/train.py:150)

    for _ in range(args.warmup):
        step()
    _sync(dev)
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        step()
        _sync(dev)
        times.append(time.perf_counter() - t0)
    ms = 1e3 * float(np.mean(times))
    print(json.dumps(dict(cmd='bench', device=dev, amp=args.amp, B=args.B, T=args.T, L=args.L, steps=args.steps, ms_per_step=ms,
                          utt_per_s=args.B / (ms / 1e3), threads=torch.get_num_threads(), dropout=bool(args.dropout),
                          torch=torch.__version__)), flush=True)


def cmd_greedy(args, rh):
    import numpy as np
    import torch
    from oracle import golden_util as gu
    cfg = gu.get_config(args.config)
    dev = args.device
    if dev == 'cpu':
        torch.set_num_threads(args.threads or os.cpu_count())
    torch.manual_seed(11785)
    model = _model(rh, cfg, dev, train=False)
    x, lx, _ = gu.make_inputs(4242, args.B, args.T, 4)
    x, lx = torch.from_numpy(x).to(dev), torch.from_numpy(lx)
    steps = cfg['speller_configs']['CHR_MAX_STEPS']

    def decode():
        with torch.inference_mode(), _autocast(dev, args.amp):
            lg, _ = model(x, lx)
        return lg.argmax(-1).cpu()

    for _ in range(args.warmup):
        decode()
    _sync(dev)
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        decode()
        _sync(dev)
        times.append(time.perf_counter() - t0)
    ms = 1e3 * float(np.mean(times))
    print(json.dumps(dict(cmd='greedy', device=dev, amp=args.amp, B=args.B, T=args.T, steps_decoded=steps, ms_per_batch=ms,
                          chars_per_s=args.B * steps / (ms / 1e3))), flush=True)


def cmd_fixture(args, rh):
    import numpy as np
    import torch
    from oracle import golden_util as gu
    g = np.load(os.path.join(ROOT, 'tests', 'golden', args.name + '.npz'), allow_pickle=False)
    dev = args.device
    if dev != 'cpu':
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
    d = [float(v) for v in g['dropout']] if 'dropout' in g.files else [0, 0, 0, 0]
    over = dict(init_dropout=d[0], mid_dropout=d[1], final_dropout=d[2], dec_lstm_dropout=d[3]) if any(d) else {}
    if 'max_steps' in g.files:
        over['CHR_MAX_STEPS'] = int(g['max_steps'])
    cfg = gu.get_config(str(g['cfg_name']), **over)
    train = 'tf_rate' in g.files
    sd = gu.make_state_dict(cfg, int(g['seed']), scale=float(g['scale']) if 'scale' in g.files else 1.0)
    model = _model(rh, cfg, dev, train=train, sd=sd)
    x, lx = torch.from_numpy(g['x']).to(dev), torch.from_numpy(g['lx'])
    out = {}
    if train:
        y = torch.from_numpy(g['y']).to(dev)
        locked = [torch.from_numpy(g[f'locked_mask_{i}']) for i in range(int(g['n_locked']))]
        drops = [torch.from_numpy(g[f'drop_mask_{i}']) for i in range(int(g['n_drops']))]
        with rh.Replayer([float(c) for c in g['coins']], locked, drops), _autocast(dev, args.amp):
            logits, att = model(x, lx, y, float(g['tf_rate']), bool(g['init_force']))
        out['logits'] = logits.detach().float().cpu().numpy()
    else:
        with torch.inference_mode(), _autocast(dev, args.amp):
            logits, att = model(x, lx)
        out['logits'] = logits.float().cpu().numpy()
        out['chars'] = out['logits'].argmax(-1)
    np.savez_compressed(args.out, **out)
    err = float(np.abs(out['logits'] - g['logits']).max())
    print(json.dumps(dict(cmd='fixture', name=args.name, device=dev, amp=args.amp, logits_abs_vs_cpu_fp32_reference=err, out=args.out)), flush=True)


def cmd_trainer(args, rh):
    """The reference's own Trainer (src/train.py) and infer_one_checkpoint (src/infer.py) on synthetic loaders."""
    import numpy as np
    import pandas as pd
    import torch
    from oracle import golden_util as gu
    import src.train as ref_train                     # the reference's driver in both modes
    import src.infer as ref_infer
    from src.utils import cfgClass
    import src.models as models
    assert os.path.abspath(os.path.dirname(ref_train.__file__)).startswith(os.path.abspath(rh.reference_root())), ref_train.__file__
    shim_active = 'las_b200' in sys.modules
    assert shim_active == bool(args.shim), (shim_active, args.shim, models.__file__)
    dev = args.device
    cfg = gu.get_config(args.config, init_dropout=0.3, mid_dropout=0.3, final_dropout=0.35, dec_lstm_dropout=0.3)
    cfg['speller_configs']['CHR_MAX_STEPS'] = args.max_steps
    tmp = tempfile.mkdtemp(prefix='las_trainer_')
    trn = dict(seed=11785, epochs=1, batch_size=args.B, accu_grad=1, grad_norm=5.0, eval_ld_interval=1, init_force=False, tf_rate=0.5,
               max_savings=3, use_specaug=False, wandb=dict(use=False, configs={}), finetune=dict(use=False, reinit_lr=False, checkpoint=''),
               model=dict(tag='base-LAS', configs=cfg),
               optimizer=dict(name='adamw', configs=dict(lr=5e-4, weight_decay=5e-6, amsgrad=True)), scaler=dict(use=args.amp != 'none'),
               batch_scheduler=dict(use=False, configs={}), epoch_scheduler=dict(use=False), tf_rate_scheduler=dict(use=False, configs={}),
               dropout_scheduler=dict(use=False, configs={}))
    trncfgs = cfgClass(trn)
    torch.manual_seed(11785)
    model = models.ListenAttendSpell(**trncfgs.model.configs).to(dev)

    def batch(seed, B, T, L, ragged=True):
        rng = np.random.default_rng(seed)
        lx = np.sort(rng.integers(T // 2, T + 1, size=B))[::-1].copy() if ragged else np.full(B, T)
        lx[0] = T
        x, lx, y = gu.make_inputs(seed, B, T, L, lx.tolist())
        y = np.concatenate([np.zeros((B, 1), dtype=np.int64), y], axis=1)          # <sos> first, like the dataset (src/utils.py)
        ly = np.full(B, L + 1, dtype=np.int64)
        return torch.from_numpy(x), torch.from_numpy(y), torch.from_numpy(lx), torch.from_numpy(ly)

    trn_loader = [batch(1, args.B, args.T, args.L), batch(2, args.B, args.T - 40, args.L - 3)]
    dev_loader = [batch(3, args.B, args.T, args.L)]
    # the reference's train_epoch calls self.scaler.unscale_ unconditionally (src/train.py:166): a scaler object is always needed
    scaler = torch.cuda.amp.GradScaler() if (args.amp != 'none' and dev != 'cpu') else torch.amp.GradScaler(dev if dev == 'cpu' else 'cuda', enabled=False)
    vocab = ['<sos>'] + [chr(ord('A') + i) for i in range(26)] + ["'", ' ', '<eos>']
    trainer = ref_train.Trainer(model=model, vocab=vocab, trn_loader=trn_loader, dev_loader=dev_loader, trncfgs=trncfgs,
                                criterion=torch.nn.CrossEntropyLoss(reduction='none'), scaler=scaler, tf_rate=0.5, saving_dir=tmp,
                                milestone_dir=tmp, device=dev, accu_grad=1, grad_norm=5.0, eval_ld_interval=1, SOS_IDX=0, EOS_IDX=29)
    t0 = time.perf_counter()
    trn_loss, trn_ppl, att = trainer.train_epoch()
    dev_loss, dev_ppl, dev_ld = trainer.evaluate_epoch()
    _sync(dev)
    t_train = time.perf_counter() - t0
    # checkpoint -> infer_one_checkpoint (src/infer.py:36-81): state_dict round trip + greedy transcripts + csv
    os.makedirs(os.path.join(tmp, 'ckpts'), exist_ok=True)
    ckpt = os.path.join(tmp, 'ckpts', 'epoch0.pt')
    torch.save({'model_state_dict': model.state_dict()}, ckpt)
    template = os.path.join(tmp, 'template.csv')
    pd.DataFrame(dict(id=list(range(args.B)), label=[''] * args.B)).to_csv(template, index=False)
    model_cfgs = cfgClass(dict(model=dict(configs=cfg)))
    infcfgs = cfgClass(dict(use_greedy=True))
    xb, _, lxb, _ = dev_loader[0]
    preds = ref_infer.infer_one_checkpoint(model_cfgs, infcfgs, ckpt, [(xb, lxb)], 'test', template, scaler, dev, vocab, 0, 29)
    print(json.dumps(dict(cmd='trainer', shim=bool(args.shim), device=dev, amp=args.amp, models_file=models.__file__,
                          train_file=ref_train.__file__, trn_loss=float(trn_loss), trn_ppl=float(trn_ppl), dev_loss=float(dev_loss),
                          dev_ld=float(dev_ld), att_shape=list(att.shape), n_preds=len(preds), pred0=preds[0][:40], seconds=t_train,
                          finite=bool(np.isfinite(trn_loss) and np.isfinite(dev_loss)))), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('cmd', choices=['bench', 'greedy', 'fixture', 'trainer'])
    ap.add_argument('--device', default='cpu')
    ap.add_argument('--amp', default='none', choices=['none', 'bf16', 'fp16'])
    ap.add_argument('--config', default='best')
    ap.add_argument('--B', type=int, default=8)
    ap.add_argument('--T', type=int, default=1600)
    ap.add_argument('--L', type=int, default=300)
    ap.add_argument('--steps', type=int, default=1)
    ap.add_argument('--warmup', type=int, default=0)
    ap.add_argument('--threads', type=int, default=0)
    ap.add_argument('--dropout', type=int, default=1)
    ap.add_argument('--name', default='best_train_T1600_L300')
    ap.add_argument('--out', default='/tmp/ref_fixture_out.npz')
    ap.add_argument('--shim', type=int, default=0)
    ap.add_argument('--max-steps', type=int, default=40)
    args = ap.parse_args()
    rh = _setup_path(bool(args.shim) and args.cmd == 'trainer')
    {'bench': cmd_bench, 'greedy': cmd_greedy, 'fixture': cmd_fixture, 'trainer': cmd_trainer}[args.cmd](args, rh)


if __name__ == '__main__':
    main()
