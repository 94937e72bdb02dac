"""
ORACLE -- TEST INFRASTRUCTURE ONLY.

Generates tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference, CPU, fp32)
on seeded weights and inputs from oracle/golden_util.py.  Runs only in the authoring container (the GPU box
has no /root/reference); the fixtures it writes are committed and are what pins the oracle and the CUDA path.

    python oracle/make_golden.py            # rewrites the small fixtures
    python oracle/make_golden.py --large    # ... and the two fixtures at the benchmarked lengths (about 3 minutes)
    python oracle/make_golden.py best_greedy_T3000      # only the named large fixture(s)

Stub modules: torchsummaryX / Levenshtein / seaborn / matplotlib are imported by the reference for
non-numerical purposes and are absent from this image.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import golden_util as gu            # noqa: E402
from oracle.las_oracle import levenshtein, idx_to_str, VOCAB   # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')


from oracle.ref_harness import Recorder, Replayer, import_reference, to_double as _to_double   # noqa: E402


def build_ref_model(ref_models, cfg, sd_np):
    import copy
    c = copy.deepcopy(cfg)
    model = ref_models.ListenAttendSpell(**c)
    ref_sd = model.state_dict()
    # the state_dict contract: same keys, same shapes
    assert set(ref_sd.keys()) == set(sd_np.keys()), (set(ref_sd) ^ set(sd_np))
    for k, v in ref_sd.items():
        assert tuple(v.shape) == tuple(sd_np[k].shape), (k, v.shape, sd_np[k].shape)
    model.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in sd_np.items()})
    assert model.spell.cls.weight is model.spell.char_emb.weight
    return model


def train_case(ref_models, name, cfg_name, seed, B, T, L, lx, ly, tf_rate, dropout=None, init_force=False,
               grads_full=True, grads_sample=0, with_fp64=False):
    over = {}
    if dropout:
        over = dict(init_dropout=dropout[0], mid_dropout=dropout[1], final_dropout=dropout[2],
                    dec_lstm_dropout=dropout[3])
    cfg = gu.get_config(cfg_name, **over)
    sd = gu.make_state_dict(cfg, seed)
    x, lx, y = gu.make_inputs(seed + 1, B, T, L, lx)
    ly = np.asarray(ly if ly is not None else [L] * B, dtype=np.int64)
    model = build_ref_model(ref_models, cfg, sd).train()
    torch.manual_seed(seed)
    with Recorder() as rec:
        logits, att = model(torch.from_numpy(x), torch.from_numpy(lx), torch.from_numpy(y), tf_rate, init_force)
    # caller-side masked CE, src/train.py:117-136 (y here is already the post-<sos> view)
    crit = torch.nn.CrossEntropyLoss(reduction='none')
    V = logits.shape[-1]
    ymask = (torch.arange(L).unsqueeze(0) < torch.from_numpy(ly).unsqueeze(1)).flatten().to(torch.int)
    loss = (crit(logits.view(-1, V), torch.from_numpy(y).view(-1)) * ymask).sum() / ymask.sum()
    loss.backward()
    out = dict(cfg_name=cfg_name, seed=seed, x=x, lx=lx, y=y, ly=ly, tf_rate=tf_rate, init_force=init_force,
               logits=logits.detach().numpy(), att=att.numpy(), loss=loss.item(),
               coins=np.asarray(rec.coins, dtype=np.float64),
               dropout=np.asarray(dropout if dropout else [0, 0, 0, 0], dtype=np.float64))
    for i, m in enumerate(rec.locked):
        out[f'locked_mask_{i}'] = m.numpy()
    for i, m in enumerate(rec.drops):
        out[f'drop_mask_{i}'] = m.numpy()
    out['n_locked'] = len(rec.locked)
    out['n_drops'] = len(rec.drops)
    nograd = []
    for k, p in model.named_parameters():
        if p.grad is None:
            nograd.append(k)
            continue
        g = p.grad.numpy()
        if grads_full:
            out['grad.' + k] = g
        elif grads_sample:
            # large configs: every gradient's norm + a strided sample of its entries (<= grads_sample values, fixed stride)
            stride = max(1, -(-g.size // grads_sample))
            out['gradsample.' + k] = g.reshape(-1)[::stride].copy()
            out['gradabsmax.' + k] = float(np.abs(g).max())
        out['gradnorm.' + k] = float(np.linalg.norm(g.astype(np.float64)))
    out['nograd'] = np.asarray(nograd)
    if with_fp64:
        # the same reference module in float64 with the recorded masks replayed: what the fp32 reference itself is off by
        m64 = _to_double(build_ref_model(ref_models, cfg, sd)).train()
        with Replayer.from_recorder(rec):
            l64, _ = m64(torch.from_numpy(x).double(), torch.from_numpy(lx), torch.from_numpy(y), tf_rate, init_force)
        loss64 = (crit(l64.view(-1, V), torch.from_numpy(y).view(-1)) * ymask).sum() / ymask.sum()
        loss64.backward()
        out['logits64'] = l64.detach().numpy()
        out['loss64'] = loss64.item()
        for k, p in m64.named_parameters():
            if p.grad is None:
                continue
            g = p.grad.numpy()
            stride = max(1, -(-g.size // grads_sample))
            out['gradsample64.' + k] = g.reshape(-1)[::stride].copy()
            out['gradnorm64.' + k] = float(np.linalg.norm(g))
        print(f'{name}: reference fp32 vs its own float64 run: logits max abs diff {np.abs(out["logits"] - out["logits64"]).max():.3e} '
              f'(max |logit| {np.abs(out["logits64"]).max():.2f}), loss diff {abs(loss.item() - loss64.item()):.3e}')
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **out)
    print(f'{name}: loss={loss.item():.6f} logits={tuple(logits.shape)} att={tuple(att.shape)} '
          f'coins={len(rec.coins)} locked={len(rec.locked)} drops={len(rec.drops)} nograd={nograd}')


def greedy_case(ref_models, name, cfg_name, seed, B, T, lx, max_steps=None, scale=1.0, with_fp64=False):
    over = {} if max_steps is None else dict(CHR_MAX_STEPS=max_steps)
    cfg = gu.get_config(cfg_name, **over)
    sd = gu.make_state_dict(cfg, seed, scale=scale)
    x, lx, y = gu.make_inputs(seed + 1, B, T, 8, lx)
    model = build_ref_model(ref_models, cfg, sd).eval()
    with torch.no_grad():
        logits, att = model(torch.from_numpy(x), torch.from_numpy(lx))
    chars = logits.argmax(-1).numpy()
    strs = [idx_to_str(c, VOCAB, 0, 29) for c in chars]
    gold = [idx_to_str(r, VOCAB, 0, 29) for r in y]
    ld = [levenshtein(s, g) for s, g in zip(strs, gold)]
    extra = {}
    if with_fp64:
        # conditioning of the fixture: the same module in float64 (must give the same transcript) and the smallest top-1 / top-2
        # logit gap per utterance -- an implementation whose logits are off by more than that gap may legitimately flip a step
        m64 = _to_double(model)
        with torch.no_grad():
            l64, _ = m64(torch.from_numpy(x).double(), torch.from_numpy(lx))
        assert np.array_equal(l64.argmax(-1).numpy(), chars), 'fixture is ill-conditioned: fp32 and float64 transcripts differ'
        top2 = torch.topk(l64, 2, -1).values
        extra = dict(logits64=l64.numpy(), margin_min=(top2[..., 0] - top2[..., 1]).min(dim=1).values.numpy())
        print(f'{name}: reference fp32 vs its own float64 run: logits max abs diff {np.abs(logits.numpy() - l64.numpy()).max():.3e}, '
              f'min top-1/top-2 gaps {extra["margin_min"]}')
    np.savez_compressed(os.path.join(OUT, name + '.npz'), cfg_name=cfg_name, seed=seed, scale=scale, x=x, lx=lx, y=y,
                        max_steps=cfg['speller_configs']['CHR_MAX_STEPS'], logits=logits.numpy(), att=att.numpy(),
                        chars=chars, transcripts=np.asarray(strs), ld=np.asarray(ld, dtype=np.int64), **extra)
    print(f'{name}: logits={tuple(logits.shape)} att={tuple(att.shape)} transcripts={strs[:2]} ld={ld}')


def optimizer_case(name, seed):
    """src/train.py:165-183 sequence with the real torch classes: GradScaler.unscale_ -> clip_grad_norm_(5.0) ->
    scaler.step(AdamW amsgrad) -> scaler.update -> zero_grad.  Step 3 carries an inf (skip path)."""
    rng = np.random.default_rng(seed)
    shapes = [(7, 5), (33,), (4, 3, 2), (129,), (2, 2)]
    p0 = [rng.standard_normal(s).astype(np.float32) for s in shapes]
    params = [torch.nn.Parameter(torch.from_numpy(a.copy())) for a in p0]
    opt = torch.optim.AdamW(params, lr=5e-4, weight_decay=5e-6, amsgrad=True)
    scaler = torch.amp.GradScaler('cpu', init_scale=65536.0)
    out = {f'p0_{i}': a for i, a in enumerate(p0)}
    n_steps = 6
    for s in range(n_steps):
        scale = scaler.get_scale()
        out[f'scale_{s}'] = scale
        for i, p in enumerate(params):
            if i == 4:          # a parameter that never receives a grad (like spell.attention.final_map)
                continue
            g = (rng.standard_normal(shapes[i]) * (3.0 if s % 2 == 0 else 0.01)).astype(np.float32)
            if s == 3 and i == 1:
                g[5] = np.inf
            out[f'g_{s}_{i}'] = g                    # UNSCALED gradient
            p.grad = torch.from_numpy(g.copy()) * scale
        # the scaler must have seen a scale() call to be "enabled + initialised"
        scaler.scale(torch.zeros(1))
        scaler.unscale_(opt)
        torch.nn.utils.clip_grad_norm_(params, 5.0)
        scaler.step(opt)
        scaler.update()
        opt.zero_grad()
        for i, p in enumerate(params):
            out[f'p_{s}_{i}'] = p.detach().numpy().copy()
    out['final_scale'] = scaler.get_scale()
    out['n_steps'] = n_steps
    out['n_params'] = len(shapes)
    st = opt.state_dict()['state']
    for i in st:
        for k in ('exp_avg', 'exp_avg_sq', 'max_exp_avg_sq'):
            out[f'state_{i}_{k}'] = st[i][k].numpy()
        out[f'state_{i}_step'] = float(st[i]['step'])
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **out)
    print(f'{name}: final scale {scaler.get_scale()} steps {[float(st[i]["step"]) for i in st]}')


def rewriter_case(name, cfg_name, seed, B, Tx, L, lx, train=True, scale=1.0):
    """src/lmtrain.py Rewriter (unmodified reference class): one train step (masked CE, backward) or one greedy decode."""
    import src.lmtrain as ref_lm
    cfg = gu.get_rewriter_config(cfg_name)
    sd = gu.make_rewriter_state_dict(cfg, seed, scale)
    model = ref_lm.Rewriter(**cfg)
    ref_sd = model.state_dict()
    assert set(ref_sd.keys()) == set(sd.keys()), set(ref_sd) ^ set(sd)
    for k, v in ref_sd.items():
        assert tuple(v.shape) == tuple(sd[k].shape), (k, v.shape, sd[k].shape)
    assert [k for k, _ in model.named_parameters()] == [k for k, _ in gu.rewriter_state_dict_shapes(cfg)]
    model.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in sd.items()})
    x, lx, y = gu.make_token_inputs(seed + 1, B, Tx, L, lx)
    out = dict(cfg_name=cfg_name, seed=seed, scale=scale, x=x, lx=lx, y=y)
    if train:
        model.train()
        torch.manual_seed(seed)
        with Recorder() as rec:
            logits, att = model(torch.from_numpy(x), torch.from_numpy(lx), torch.from_numpy(y), 0.5)
        loss = torch.nn.functional.cross_entropy(logits.reshape(-1, logits.shape[-1]), torch.from_numpy(y).reshape(-1))
        loss.backward()
        out.update(logits=logits.detach().numpy(), att=att.numpy(), loss=loss.item(), n_coins=len(rec.coins))
        nograd = []
        for k, p in model.named_parameters():
            if p.grad is None:
                nograd.append(k)
                continue
            out['grad.' + k] = p.grad.numpy()
        out['nograd'] = np.asarray(nograd)
        print(f'{name}: loss={loss.item():.6f} logits={tuple(logits.shape)} att={tuple(att.shape)} coins={len(rec.coins)} nograd={nograd}')
    else:
        model.eval()
        with torch.no_grad():
            logits, att = model(torch.from_numpy(x), torch.from_numpy(lx))
        chars = logits.argmax(-1).numpy()
        out.update(logits=logits.numpy(), att=att.numpy(), chars=chars)
        print(f'{name}: logits={tuple(logits.shape)} chars[0]={chars[0][:12]}')
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **out)


def collate_case(name, seed, lens, F=15, specaug=True):
    """datasetTrainDev.collate_fn of the unmodified reference (src/utils.py:95-128) on a ragged synthetic batch.  The class
    constructor reads .npy directories, so the instance is built without it and given exactly the attributes collate_fn uses."""
    import src.utils as ref_utils
    import torchaudio.transforms as tat
    ds = object.__new__(ref_utils.datasetTrainDev)
    ds.useSpecAug = specaug
    ds.freq_masker = tat.FrequencyMasking(6)            # src/utils.py:83-84
    ds.time_masker = tat.TimeMasking(200)
    rng = np.random.default_rng(seed)
    mf = [rng.standard_normal((n, F)).astype(np.float32) for n in lens]
    tr = [rng.integers(1, 29, size=int(rng.integers(3, 12))).astype(np.int64) for _ in lens]
    torch.manual_seed(seed)
    x, y, lx, ly = ds.collate_fn([(torch.from_numpy(a), torch.from_numpy(b)) for a, b in zip(mf, tr)])
    out = dict(seed=seed, specaug=specaug, n=len(lens), x=x.numpy(), y=y.numpy(), lx=lx.numpy(), ly=ly.numpy())
    for i, (a, b) in enumerate(zip(mf, tr)):
        out[f'mfcc_{i}'] = a
        out[f'trans_{i}'] = b
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **out)
    print(f'{name}: x={tuple(x.shape)} y={tuple(y.shape)} lx={lx.tolist()} masked_frac={(x == 0).float().mean():.3f}')


def main_large(ref_models, only):
    """Fixtures at BASELINE.json's benchmarked lengths (configs[1]: T=1600, L=300; configs[3]: T=3000, 600 greedy steps) with the
    best config of config/sample-attention.yml:42-68 INCLUDING its dropouts (0.3 / 0.3 / 0.35, decoder 0.3; masks recorded).  Batch
    of 3-4 ragged utterances so that the unmodified reference finishes in about a minute on the CPU."""
    if not only or 'best_train_T1600_L300' in only:
        train_case(ref_models, 'best_train_T1600_L300', 'best', 1111, B=3, T=1600, L=300, lx=[1600, 1433, 1197], ly=[300, 257, 281],
                   tf_rate=1.0, dropout=(0.3, 0.3, 0.35, 0.3), grads_full=False, grads_sample=4096, with_fp64=True)
    if not only or 'best_greedy_T3000' in only:
        # weights x2.75 / seed 1500: between the fixed-point regime (smaller scale: 'HHHH...' for 600 steps) and the chaotic one
        # (larger scale: the reference's own fp32 and float64 transcripts differ) -- every utterance has 30-350 token changes,
        # fp32 == float64 transcript
        greedy_case(ref_models, 'best_greedy_T3000', 'best', 1500, B=4, T=3000, lx=[3000, 2871, 2500, 1999], scale=2.75,
                    with_fp64=True)


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    ref_models = import_reference()
    only = [a for a in sys.argv[1:] if not a.startswith('-')]
    if '--large' in sys.argv[1:] or only:
        main_large(ref_models, only)
        if only or '--large-only' in sys.argv[1:]:
            return
    # (1) micro, ragged + unsorted lengths, odd at every pyramid level, teacher forcing 1.0, no dropout
    train_case(ref_models, 'micro_train_tf1', 'micro', 101, B=3, T=37, L=7, lx=[19, 37, 30], ly=[7, 5, 6], tf_rate=1.0)
    # (2) micro, max(lx) < padded T, a row of length exactly 8 (-> enc_len 1), tf 0.5 (coins recorded)
    train_case(ref_models, 'micro_train_tf05', 'micro', 202, B=4, T=50, L=9, lx=[45, 8, 33, 26], ly=None, tf_rate=0.5)
    # (3) micro with every dropout on (masks recorded)
    train_case(ref_models, 'micro_train_dropout', 'micro', 303, B=3, T=41, L=6, lx=[41, 40, 17], ly=None,
               tf_rate=1.0, dropout=(0.3, 0.3, 0.35, 0.3))
    # (4) micro with the init_force diagonal prior
    train_case(ref_models, 'micro_train_initforce', 'micro', 404, B=2, T=48, L=8, lx=[48, 40], ly=None, tf_rate=1.0,
               init_force=True)
    # (5) tiny = BASELINE configs[0]: B=4, T=400, hid 128, 1 pLSTM, one teacher-forced step
    train_case(ref_models, 'tiny_train_tf1', 'tiny', 505, B=4, T=400, L=30, lx=[400, 380, 333, 251], ly=None,
               tf_rate=1.0, grads_full=False)
    # (6) greedy decode, micro and tiny (weights scaled up so transcripts are not degenerate)
    greedy_case(ref_models, 'micro_greedy', 'micro', 606, B=3, T=37, lx=[37, 21, 30], scale=3.0)
    greedy_case(ref_models, 'tiny_greedy', 'tiny', 707, B=4, T=400, lx=[400, 380, 333, 251], scale=2.0)
    # (7) optimizer
    optimizer_case('optimizer_adamw_amsgrad', 808)
    # (8) Rewriter (src/lmtrain.py): ragged token inputs, 4 heads; train step (tf_rate 0.5: the coin is drawn, never used) + greedy
    rewriter_case('rewriter_train', 'rw_micro', 909, B=3, Tx=11, L=6, lx=[11, 7, 9], train=True)
    rewriter_case('rewriter_greedy', 'rw_micro', 910, B=3, Tx=13, L=6, lx=[13, 5, 10], train=False, scale=3.0)
    # (9) loader collate + SpecAugment (src/utils.py:95-128): unsorted ragged lengths incl. ties, T < 200 and T > 200
    collate_case('collate_specaug_short', 1001, lens=[50, 37, 64, 12, 64])
    collate_case('collate_specaug_long', 1002, lens=[310, 250, 333])
    collate_case('collate_plain', 1003, lens=[9, 30, 21], specaug=False)


if __name__ == '__main__':
    main()
