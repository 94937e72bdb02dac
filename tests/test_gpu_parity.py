"""GPU parity tests (run on the B200 box with `-m gpu`): the CUDA path, called through the C ABI, against
  (1) the golden fixtures produced by the unmodified reference, and
  (2) the oracle restatement on seeded inputs at other shapes,
plus size-independent properties at larger sizes.  Tolerance: north_star's fp32 bar, 1e-4 relative, written below."""
import numpy as np
import pytest
import torch

from helpers import (gu, orc, load_golden, fixture_cfg, fixture_masks, oracle_train_from_fixture, rel_err, grad_floor)

pytestmark = pytest.mark.gpu
TOL = 1e-4          # fp32 logits and gradients: <= 1e-4 relative error (BASELINE.json north_star)
DEV = 'cuda:0'


def _model(cfg, sd, train=True):
    from las_b200.models import ListenAttendSpell
    m = ListenAttendSpell(**cfg).to(DEV)
    m.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd.items()})
    return m.train() if train else m.eval()


def _masked_ce(logits, y, ly):
    B, L, V = logits.shape
    crit = torch.nn.CrossEntropyLoss(reduction='none')
    mask = (torch.arange(L, device=logits.device).unsqueeze(0) < torch.as_tensor(ly, device=logits.device).unsqueeze(1)).flatten().to(torch.int)
    return (crit(logits.view(-1, V), y.view(-1)) * mask).sum() / mask.sum()


# ----------------------------------------------------------------------------------------------------------------------
# building blocks
# ----------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('M,N,K', [(1, 1, 1), (96, 2048, 768), (300, 30, 512), (257, 129, 65), (1000, 64, 15), (4096, 512, 300),
                                   (28800, 30, 512),        # the tied classifier over all decoder steps: the tall 128 x 32 tile
                                   (19001, 32, 70),         # same tile, N a multiple of 4: its 128-bit epilogue, ragged last row tile
                                   (28800, 512, 30)])       # dQC = dlogits . emb: all epilogue (128-bit stores)
def test_gemm_f32_vs_torch(M, N, K):
    from las_b200 import functional as LF
    g = torch.Generator(device='cpu').manual_seed(M * 7 + N)
    A = torch.randn(M, K, generator=g).to(DEV)
    Bm = torch.randn(N, K, generator=g).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    C = torch.empty(M, N, device=DEV)
    LF.gemm_raw(A, Bm, C, M, N, K, am=(0, K, 0), ak=(0, 1, 0), bk=(0, 1, 0), bn=K, cm=(0, N, 0), bias1=bias)
    ref = (A.double() @ Bm.double().t() + bias.double()).float()
    assert rel_err(C.cpu().numpy(), ref.cpu().numpy()) < 1e-5
    # weight-gradient form: C2[n][k] = sum_m A[m][k] * D[m][n]  (both operands strided along the reduction)
    D = torch.randn(M, N, generator=g).to(DEV)
    C2 = torch.empty(N, K, device=DEV)
    LF.gemm_raw(D, A, C2, N, K, M, am=(0, 1, 0), ak=(0, N, 0), bk=(0, K, 0), bn=1, cm=(0, K, 0))
    ref2 = (D.double().t() @ A.double()).float()
    assert rel_err(C2.cpu().numpy(), ref2.cpu().numpy()) < 1e-5
    # accumulate form (beta = 1, alpha = 0.5, both biases) into a padded row pitch, and into a pitch that is NOT a multiple of four
    # floats (scalar epilogue): every epilogue branch sees the same arithmetic
    for pitch in (N + 4 - N % 4 if N % 4 else N + 8, N + 1):
        C3 = torch.randn(M, pitch, generator=g).to(DEV)
        ref3 = C3.clone()
        ref3[:, :N] = (0.5 * (A.double() @ Bm.double().t()) + C3[:, :N].double() + 2 * bias.double()).float()
        LF.gemm_raw(A, Bm, C3, M, N, K, am=(0, K, 0), ak=(0, 1, 0), bk=(0, 1, 0), bn=K, cm=(0, pitch, 0), bias1=bias, bias2=bias,
                    alpha=0.5, beta=1.0)
        assert rel_err(C3.cpu().numpy(), ref3.cpu().numpy()) < 1e-5        # columns past N untouched


def test_gemm_two_level_index_is_the_pyramid_reshape():
    """(B,T,D) -> drop odd frame -> (B,T//2,2D) done purely by addressing (reference src/modules.py:171-185)."""
    from las_b200 import functional as LF
    B, T, D, N = 3, 9, 8, 16
    x = torch.randn(B, T, D, device=DEV)
    W = torch.randn(N, 2 * D, device=DEV)
    Tp = T // 2
    out = torch.empty(B, Tp, N, device=DEV)
    LF.gemm_raw(x, W, out, B * Tp, N, 2 * D, am=(T * D, 2 * D, Tp), ak=(0, 1, 0), bk=(0, 1, 0), bn=2 * D, cm=(0, N, 0))
    ref = x[:, :2 * Tp].reshape(B, Tp, 2 * D) @ W.t()
    assert rel_err(out.cpu().numpy(), ref.cpu().numpy()) < 1e-5


@pytest.mark.parametrize('H,Din,B,T,lens,pyramid,use_mask', [
    (32, 15, 3, 11, [11, 4, 7], False, False),
    (32, 24, 4, 13, [5, 6, 1, 3], True, True),        # odd T, a row of length 1, unsorted, locked dropout
    (128, 15, 5, 40, [40, 33, 40, 2, 17], False, True),
    (64, 32, 100, 6, None, True, False),              # more rows than one 96-row chunk
])
def test_lstm_layer_vs_oracle(H, Din, B, T, lens, pyramid, use_mask):
    """One BiLSTM layer with PackedSequence semantics (+ pyramid, + locked dropout): output, and gradients of the input
    and of all eight parameter tensors."""
    from las_b200 import functional as LF
    rng = np.random.default_rng(H + B + T)
    D = Din // 2 if pyramid else Din
    Tin = T
    x = torch.from_numpy(rng.standard_normal((B, Tin, D)).astype(np.float32))
    if lens is None:
        lens = list(rng.integers(1, T // 2 + 1 if pyramid else T + 1, size=B))
        lens[0] = T // 2 if pyramid else T
    lens = [int(v) for v in lens]
    To = max(lens)
    k = 1 / np.sqrt(H)
    names = ['weight_ih_l0', 'weight_hh_l0', 'bias_ih_l0', 'bias_hh_l0']
    shapes = [(4 * H, Din), (4 * H, H), (4 * H,), (4 * H,)]
    p = {}
    for suf in ['', '_reverse']:
        for n, s in zip(names, shapes):
            p['l.' + n + suf] = torch.from_numpy(rng.uniform(-k, k, size=s).astype(np.float32))
    mask = None
    if use_mask:
        mask = torch.from_numpy((rng.random((B, 1, 2 * H)) > 0.3).astype(np.float32) / 0.7)
    wout = torch.from_numpy(rng.standard_normal((B, To, 2 * H)).astype(np.float32))
    # oracle
    po = {k_: v.clone().requires_grad_(True) for k_, v in p.items()}
    xo = x.clone().requires_grad_(True)
    xin = xo[:, :2 * (Tin // 2)].reshape(B, Tin // 2, 2 * D) if pyramid else xo
    yo = orc.bilstm_layer(xin, lens, po, 'l.')
    if mask is not None:
        yo = yo * mask
    (yo * wout).sum().backward()
    # CUDA path
    pc = {k_: v.clone().to(DEV).requires_grad_(True) for k_, v in p.items()}
    xc = x.clone().to(DEV).requires_grad_(True)
    ws = [pc['l.' + n + suf] for suf in ['', '_reverse'] for n in names]
    lens_dev = torch.tensor(lens, dtype=torch.int32, device=DEV)
    yc = LF.lstm_layer(xc, lens_dev, To, pyramid, mask.to(DEV) if mask is not None else None, ws)
    assert tuple(yc.shape) == tuple(yo.shape)
    (yc * wout.to(DEV)).sum().backward()
    assert rel_err(yc.detach().cpu().numpy(), yo.detach().numpy()) < TOL
    # exact zeros beyond each row's length (pad_packed_sequence)
    for b, l in enumerate(lens):
        assert float(yc[b, l:].abs().max()) == 0.0 if l < To else True
    assert rel_err(xc.grad.cpu().numpy(), xo.grad.numpy()) < TOL
    for k_ in p:
        assert rel_err(pc[k_].grad.cpu().numpy(), po[k_].grad.numpy()) < TOL, k_


@pytest.mark.parametrize('B,T,P,heads,lens', [(3, 5, 16, 1, [5, 1, 3]), (4, 200, 256, 1, [200, 187, 31, 100]),
                                              (2, 37, 64, 4, [37, 20]), (5, 375, 128, 1, None)])
@pytest.mark.parametrize('split', [None, 0, 1, 2, 3, 8])
def test_attention_step_vs_oracle(B, T, P, heads, lens, split, monkeypatch):
    """split = CTAs per (batch row, head) of the single-pass T-split kernel (None: the library's own choice; 0: the
    two-phase one-CTA-per-row kernel).  Rows shorter than the split leave some CTAs of the cluster without work."""
    from las_b200 import functional as LF
    if split is None:
        monkeypatch.delenv('LAS_ATTN_SPLIT', raising=False)
    else:
        monkeypatch.setenv('LAS_ATTN_SPLIT', str(split))
    rng = np.random.default_rng(B * T + P)
    if lens is None:
        lens = [T] + [int(v) for v in rng.integers(1, T + 1, size=B - 1)]
    q = torch.from_numpy(rng.standard_normal((B, P)).astype(np.float32) * 0.3)
    K = torch.from_numpy(rng.standard_normal((B, T, P)).astype(np.float32) * 0.3)
    V = torch.from_numpy(rng.standard_normal((B, T, P)).astype(np.float32))
    wctx = torch.from_numpy(rng.standard_normal((B, P)).astype(np.float32))
    d = P // heads
    # oracle: attention_step without the query projection (identity query_map)
    po = {'a.query_map.weight': torch.eye(P), 'a.query_map.bias': torch.zeros(P)}
    qo, Ko, Vo = q.clone().requires_grad_(True), K.clone().requires_grad_(True), V.clone().requires_grad_(True)
    keys = Ko.view(B, T, heads, d).permute(0, 2, 1, 3)
    vals = Vo.view(B, T, heads, d).permute(0, 2, 1, 3)
    pad = torch.arange(T).unsqueeze(0) >= torch.tensor(lens).unsqueeze(1)
    ctx_o, w_o, _ = orc.attention_step(po, qo, keys, vals, pad, heads, None, 'a.')
    (ctx_o * wctx).sum().backward()
    qc, Kc, Vc = (t.clone().to(DEV).requires_grad_(True) for t in (q, K, V))
    ctx_c, w_c = LF.attn_step(qc, Kc, Vc, torch.tensor(lens, dtype=torch.int32, device=DEV), heads)
    (ctx_c * wctx.to(DEV)).sum().backward()
    assert rel_err(ctx_c.detach().cpu().numpy(), ctx_o.detach().numpy()) < TOL
    assert np.abs(w_c.cpu().numpy() - w_o.detach().numpy()).max() < 1e-5
    for b, l in enumerate(lens):          # masked weights are exactly zero (src/models.py:174-175)
        if l < T:
            assert float(w_c[b, :, l:].abs().max()) == 0.0
    assert rel_err(qc.grad.cpu().numpy(), qo.grad.numpy()) < TOL
    assert rel_err(Kc.grad.cpu().numpy(), Ko.grad.numpy()) < TOL
    assert rel_err(Vc.grad.cpu().numpy(), Vo.grad.numpy()) < TOL


@pytest.mark.parametrize('B,T,P,heads,lens', [(3, 23, 16, 1, [23, 9, 17]), (2, 40, 64, 4, [40, 22])])
def test_attention_step_init_force_prior_vs_oracle(B, T, P, heads, lens):
    """init_wgts_mask path of MultiheadCrossAttention.forward (src/models.py:177-181): second softmax over all T
    positions (pads included), pre-prior weights returned."""
    from las_b200 import functional as LF
    rng = np.random.default_rng(7 * B + T)
    q = torch.from_numpy(rng.standard_normal((B, P)).astype(np.float32) * 0.3)
    K = torch.from_numpy(rng.standard_normal((B, T, P)).astype(np.float32) * 0.3)
    V = torch.from_numpy(rng.standard_normal((B, T, P)).astype(np.float32))
    wctx = torch.from_numpy(rng.standard_normal((B, P)).astype(np.float32))
    prior = torch.zeros(T)
    prior[T // 4: T // 4 + T // 3] = 1.0
    d = P // heads
    po = {'a.query_map.weight': torch.eye(P), 'a.query_map.bias': torch.zeros(P)}
    qo, Ko, Vo = q.clone().requires_grad_(True), K.clone().requires_grad_(True), V.clone().requires_grad_(True)
    keys = Ko.view(B, T, heads, d).permute(0, 2, 1, 3)
    vals = Vo.view(B, T, heads, d).permute(0, 2, 1, 3)
    pad = torch.arange(T).unsqueeze(0) >= torch.tensor(lens).unsqueeze(1)
    ctx_o, w_o, _ = orc.attention_step(po, qo, keys, vals, pad, heads, prior.expand(B, heads, T), 'a.')
    (ctx_o * wctx).sum().backward()
    qc, Kc, Vc = (t.clone().to(DEV).requires_grad_(True) for t in (q, K, V))
    fmask = prior.to(DEV).expand(B * heads, T).contiguous()
    ctx_c, w_c = LF.attn_step(qc, Kc, Vc, torch.tensor(lens, dtype=torch.int32, device=DEV), heads, fmask)
    (ctx_c * wctx.to(DEV)).sum().backward()
    assert rel_err(ctx_c.detach().cpu().numpy(), ctx_o.detach().numpy()) < TOL
    assert np.abs(w_c.cpu().numpy() - w_o.detach().numpy()).max() < 1e-5
    assert rel_err(qc.grad.cpu().numpy(), qo.grad.numpy()) < TOL
    assert rel_err(Kc.grad.cpu().numpy(), Ko.grad.numpy()) < TOL
    assert rel_err(Vc.grad.cpu().numpy(), Vo.grad.numpy()) < TOL


# ----------------------------------------------------------------------------------------------------------------------
# whole model against the reference's golden fixtures
# ----------------------------------------------------------------------------------------------------------------------
TRAIN_CASES = ['micro_train_tf1', 'micro_train_tf05', 'micro_train_dropout', 'micro_train_initforce', 'tiny_train_tf1']


@pytest.mark.parametrize('name', TRAIN_CASES)
def test_las_train_step_matches_reference_golden(name):
    from las_b200.modules import set_mask_override
    g = load_golden(name)
    cfg = fixture_cfg(g)
    sd = gu.make_state_dict(cfg, int(g['seed']))
    model = _model(cfg, sd, train=True)
    nl, nd = int(g['n_locked']), int(g['n_drops'])
    locked = [torch.from_numpy(g[f'locked_mask_{i}']) for i in range(nl)] if nl else None
    drops = [torch.from_numpy(g[f'drop_mask_{i}']) for i in range(nd)] if nd else None
    set_mask_override(locked, drops, [float(c) for c in g['coins']])
    try:
        y = torch.from_numpy(g['y']).to(DEV)
        logits, att = model(torch.from_numpy(g['x']).to(DEV), torch.from_numpy(g['lx']), y, float(g['tf_rate']),
                            bool(g['init_force']))
        loss = _masked_ce(logits, y, g['ly'])
        loss.backward()
    finally:
        set_mask_override(None, None, None)
    assert rel_err(logits.detach().cpu().numpy(), g['logits']) < TOL
    assert not att.is_cuda and tuple(att.shape) == g['att'].shape          # (heads, T_enc, steps+1) CPU tensor
    assert np.abs(att.numpy() - g['att']).max() < 1e-5
    assert abs(float(loss) - float(g['loss'])) < 1e-5
    floor = grad_floor(g)
    nograd = set(str(s) for s in g['nograd'])
    for k, p in model.named_parameters():
        if k in nograd:
            assert p.grad is None, k                                       # final_map never gets a grad
            continue
        ref_norm = float(g['gradnorm.' + k])
        got = float(np.linalg.norm(p.grad.cpu().numpy().astype(np.float64)))
        assert abs(got - ref_norm) <= TOL * max(ref_norm, floor), (k, got, ref_norm)
        if ('grad.' + k) in g.files:
            assert rel_err(p.grad.cpu().numpy(), g['grad.' + k], floor) < TOL, k


@pytest.mark.parametrize('name', ['micro_greedy', 'tiny_greedy'])
def test_greedy_transcripts_identical_to_reference(name):
    g = load_golden(name)
    cfg = fixture_cfg(g)
    sd = gu.make_state_dict(cfg, int(g['seed']), scale=float(g['scale']))
    model = _model(cfg, sd, train=False)
    with torch.no_grad():
        logits, att = model(torch.from_numpy(g['x']).to(DEV), torch.from_numpy(g['lx']))
    chars = logits.argmax(-1).cpu().numpy()
    assert np.array_equal(chars, g['chars'])                               # greedy index sequences identical
    assert np.array_equal(model.spell.last_chars.t().cpu().numpy(), g['chars'])   # the fed-back argmax is the same one
    strs = [orc.idx_to_str(c, orc.VOCAB, 0, 29) for c in chars]
    assert strs == [str(s) for s in g['transcripts']]
    gold = [orc.idx_to_str(r, orc.VOCAB, 0, 29) for r in g['y']]
    assert [orc.levenshtein(a, b) for a, b in zip(strs, gold)] == g['ld'].tolist()   # equal Levenshtein distance
    assert rel_err(logits.cpu().numpy(), g['logits']) < TOL
    assert tuple(att.shape) == g['att'].shape
    assert np.abs(att.numpy() - g['att']).max() < 1e-5


# ----------------------------------------------------------------------------------------------------------------------
# whole model against the oracle at the best config's dims (small B/T so the CPU oracle finishes in seconds)
# ----------------------------------------------------------------------------------------------------------------------
def test_best_config_dims_vs_oracle():
    cfg = gu.get_config('best')
    sd = gu.make_state_dict(cfg, 11)
    B, T, L = 3, 72, 6
    x, lx, y = gu.make_inputs(12, B, T, L, [72, 49, 64])
    model = _model(cfg, sd, train=True)
    yd = torch.from_numpy(y).to(DEV)
    logits, att = model(torch.from_numpy(x).to(DEV), torch.from_numpy(lx), yd, 1.0, False)
    loss = _masked_ce(logits, yd, [L] * B)
    loss.backward()
    p = {k: torch.from_numpy(v.copy()).requires_grad_(True) for k, v in sd.items() if k != 'spell.cls.weight'}
    p['spell.cls.weight'] = p['spell.char_emb.weight']
    ol, oatt = orc.las_forward(p, torch.from_numpy(x), lx.tolist(), lstm_layers=1, plstm_layers=3, heads=1, training=True,
                               steps=L, dec_y=torch.from_numpy(y), coins=[True] * L)
    oloss = orc.masked_ce_loss(ol, torch.from_numpy(y), [L] * B)
    oloss.backward()
    assert rel_err(logits.detach().cpu().numpy(), ol.detach().numpy()) < TOL
    gmax = max(float(v.grad.abs().max()) for k, v in p.items() if v.grad is not None)
    for k, prm in model.named_parameters():
        if p[k].grad is None:
            assert prm.grad is None
            continue
        assert rel_err(prm.grad.cpu().numpy(), p[k].grad.numpy(), 1e-3 * gmax) < TOL, k


# ----------------------------------------------------------------------------------------------------------------------
# properties that hold at any size
# ----------------------------------------------------------------------------------------------------------------------
def test_rows_are_independent_and_padding_is_inert():
    """Utterances are independent end to end: a row's logits do not change when other rows / the padded
    tail change, and the result is deterministic run to run."""
    cfg = gu.get_config('tiny')
    sd = gu.make_state_dict(cfg, 21)
    model = _model(cfg, sd, train=False)
    x, lx, _ = gu.make_inputs(22, 6, 160, 4, [160, 120, 96, 160, 33, 80])
    xd = torch.from_numpy(x).to(DEV)
    with torch.no_grad():
        a, _ = model(xd, torch.from_numpy(lx))
        b, _ = model(xd, torch.from_numpy(lx))
        assert torch.equal(a, b)
        x2 = xd.clone()
        x2[1, 120:] = 123.0                       # garbage in row 1's padding
        x2[3] = torch.randn_like(x2[3])           # different utterance in row 3
        c, _ = model(x2, torch.from_numpy(lx))
        keep = [0, 1, 2, 4, 5]
        assert torch.equal(a[keep], c[keep])
        d, _ = model(xd[:3].contiguous(), torch.from_numpy(lx[:3]))     # smaller batch, same rows
        assert rel_err(d.cpu().numpy(), a[:3].cpu().numpy()) < 1e-5


def test_fused_adamw_matches_torch_golden():
    from las_b200.optim import FusedAdamW
    g = load_golden('optimizer_adamw_amsgrad')
    n, steps = int(g['n_params']), int(g['n_steps'])
    params = [torch.nn.Parameter(torch.from_numpy(g[f'p0_{i}'].copy()).to(DEV)) for i in range(n)]
    opt = FusedAdamW(params, lr=5e-4, weight_decay=5e-6, amsgrad=True)
    for s in range(steps):
        scale = float(g[f'scale_{s}'])
        for i, p in enumerate(params):
            p.grad = (torch.from_numpy(g[f'g_{s}_{i}']).to(DEV) * scale) if f'g_{s}_{i}' in g.files else None
        status = opt.step_fused(inv_scale=1.0 / scale, max_norm=5.0, sync_skip=True)
        assert bool(status[0].item() != 0) == (s == 3)          # the injected inf skips the step
        for i, p in enumerate(params):
            np.testing.assert_allclose(p.detach().cpu().numpy(), g[f'p_{s}_{i}'], rtol=2e-6, atol=1e-7)
    sd = opt.state_dict()['state']
    for i in range(n - 1):
        assert float(sd[i]['step']) == float(g[f'state_{i}_step'])
        np.testing.assert_allclose(sd[i]['max_exp_avg_sq'].cpu().numpy(), g[f'state_{i}_max_exp_avg_sq'], rtol=1e-5, atol=1e-9)
    assert 4 not in sd or float(sd[4]['step']) == 0            # the grad-less parameter was never touched


def test_decoder_graph_replays_in_a_steady_loop_and_slots_do_not_alias():
    """The Speller forward loop is one CUDA graph keyed on the descriptor's pointers; the Python wrapper stages everything in
    pooled, pointer-stable buffers.  A steady train loop must capture once and replay afterwards; two forwards that are both
    alive (no backward in between) must not share history buffers."""
    import ctypes as C
    from las_b200 import _lib
    lib = _lib.load()
    cfg = gu.get_config('micro')
    sd = gu.make_state_dict(cfg, 5)
    model = _model(cfg, sd, train=True)
    x, lx, y = gu.make_inputs(9, 3, 40, 6, [40, 25, 33])
    xd, yd = torch.from_numpy(x).to(DEV), torch.from_numpy(y).to(DEV)

    def stats():
        c, r = C.c_longlong(), C.c_longlong()
        lib.las_speller_graph_stats(C.byref(c), C.byref(r))
        return c.value, r.value

    grads = []
    for it in range(5):
        model.zero_grad(set_to_none=True)
        junk = torch.empty(1000 + 997 * it, device=DEV)          # perturb the caching allocator between steps
        logits, _ = model(xd, torch.from_numpy(lx), yd, 1.0, False)
        del junk
        logits.square().mean().backward()
        if it == 1:                # a key is captured on its first sighting, or on its second one when the cache has been missing
            c0, r0 = stats()       # for a while (non-recurring shapes are enqueued directly): steady from the third step at the latest
        grads.append(model.spell.attention.query_map.weight.grad.clone())
    c1, r1 = stats()
    assert c1 == c0, 'steady loop re-captured a decoder graph'
    assert r1 - r0 == 6            # 3 more steps x (forward loop + backward loop), all replays
    for g in grads[1:]:
        assert torch.equal(g, grads[0])
    # two live forwards: the first one's backward must still see its own history
    model.zero_grad(set_to_none=True)
    la, _ = model(xd, torch.from_numpy(lx), yd, 1.0, False)
    y2 = torch.from_numpy(np.roll(y, 1, axis=1).copy()).to(DEV)
    lb, _ = model(xd, torch.from_numpy(lx), y2, 1.0, False)
    la.square().mean().backward()
    assert torch.equal(model.spell.attention.query_map.weight.grad, grads[0])
    del lb


@pytest.mark.parametrize('B,L,V,ly,accu', [(3, 7, 30, [7, 5, 6], 1), (5, 33, 30, [33, 1, 17, 30, 8], 4), (2, 4, 70, [4, 2], 1)])
def test_fused_masked_ce_matches_trainer_loss(B, L, V, ly, accu):
    """las_b200.loss.masked_ce against the reference trainer's loss (src/train.py:117-136) and the oracle's restatement:
    loss value, perplexity and d loss / d logits."""
    from las_b200.loss import masked_ce
    rng = np.random.default_rng(B * L + V)
    logits = torch.from_numpy(rng.standard_normal((B, L, V)).astype(np.float32) * 3)
    y = torch.from_numpy(rng.integers(0, V, size=(B, L)).astype(np.int64))
    lo = logits.clone().requires_grad_(True)
    y_mask = (torch.arange(L).unsqueeze(0) < torch.tensor(ly).unsqueeze(1)).flatten().to(torch.int)
    ref = (torch.nn.CrossEntropyLoss(reduction='none')(lo.view(-1, V), y.view(-1)) * y_mask).sum() / (y_mask.sum() * accu)
    (ref * 65536.0).backward()
    if V == 30:
        assert abs(float(orc.masked_ce_loss(logits, y, ly, accu)) - float(ref)) < 1e-6
    lc = logits.clone().to(DEV).requires_grad_(True)
    loss, ppl = masked_ce(lc, y.to(DEV), torch.tensor(ly), accu)
    (loss * 65536.0).backward()
    assert abs(float(loss) - float(ref)) < 1e-5 * max(1.0, abs(float(ref)))
    assert abs(float(ppl) - float(torch.exp(ref))) < 1e-4 * float(torch.exp(ref))
    assert rel_err(lc.grad.cpu().numpy(), lo.grad.numpy()) < TOL
    l2, _ = masked_ce(lc.detach(), y.to(DEV), torch.tensor(ly), accu)
    assert float(l2) == float(loss)                      # deterministic
    # dev-eval form (src/train.py:226-232): a longer decode truncated to the target length
    longer = torch.cat([logits, torch.from_numpy(rng.standard_normal((B, 5, V)).astype(np.float32))], dim=1).to(DEV)
    l3, _ = masked_ce(longer, y.to(DEV), torch.tensor(ly), accu)
    assert float(l3) == float(loss)


def test_decoder_buffer_pool_is_bounded_over_ragged_batches():
    """The reference trainer's ragged batches give a new (T, steps) almost every batch.  The pooled decoder buffers must settle on one
    buffer set sized for the largest batch (not one set per distinct shape), survive evaluation under torch.inference_mode()
    (src/train.py:207) followed by decoding outside it (src/infer.py:56-62), and stay flat in device memory."""
    from las_b200 import functional as LF
    cfg = gu.get_config('tiny')
    sd = gu.make_state_dict(cfg, 31)
    model = _model(cfg, sd, train=True)
    rng = np.random.default_rng(0)
    shapes = [(int(rng.integers(100, 201)) // 2 * 2, int(rng.integers(5, 26))) for _ in range(40)]
    shapes[7] = (200, 25)                                    # the largest batch arrives early

    def train_step(T, L):
        x, lx, y = gu.make_inputs(T * 31 + L, 4, T, L)
        yd = torch.from_numpy(y).to(DEV)
        model.zero_grad(set_to_none=True)
        logits, _ = model(torch.from_numpy(x).to(DEV), torch.from_numpy(lx), yd, 0.5, False)
        _masked_ce(logits, yd, [L] * 4).backward()

    LF.speller_pool_clear()
    train_step(200, 25)
    torch.cuda.synchronize()
    one_set = LF.speller_pool_bytes()
    LF.speller_pool_clear()
    marks = []
    for i, (T, L) in enumerate(shapes):
        train_step(T, L)
        if i in (9, 39):
            torch.cuda.synchronize()
            marks.append((LF.speller_pool_bytes(), torch.cuda.memory_allocated()))
    assert marks[1][0] == marks[0][0], marks                 # no growth after the largest shape has been seen
    assert marks[1][0] <= 1.3 * one_set, (marks, one_set)    # one buffer set, not one per shape
    assert marks[1][1] <= marks[0][1] + (8 << 20), marks     # allocator footprint flat too
    # evaluation under inference_mode, then decoding outside it: the pooled buffers must stay writable
    model.eval()
    x, lx, _ = gu.make_inputs(5, 4, 160, 4)
    xd = torch.from_numpy(x).to(DEV)
    LF.speller_pool_clear()
    with torch.inference_mode():
        a, _ = model(xd, torch.from_numpy(lx))
    b, _ = model(xd, torch.from_numpy(lx))
    with torch.no_grad():
        c, _ = model(xd, torch.from_numpy(lx))
    assert torch.equal(a, b.detach()) and torch.equal(a, c)


def test_fused_masked_ce_out_of_range_target_is_loud():
    """nn.CrossEntropyLoss raises on a non-masked target outside [0, V) (the reference trainer's criterion); the sync-free fused loss
    turns the loss and that row's gradient into NaN instead of dropping the one-hot term silently.  Masked positions may hold anything."""
    from las_b200.loss import masked_ce
    logits = torch.randn(2, 4, 30, device=DEV, requires_grad=True)
    y = torch.tensor([[1, 2, 3, 99], [4, 77, 5, 6]], device=DEV)
    loss, _ = masked_ce(logits, y, torch.tensor([3, 4]), 1)          # row 0's bad target is masked, row 1's is not
    assert torch.isnan(loss)
    loss.backward()
    assert torch.isnan(logits.grad[1, 1]).all() and torch.isfinite(logits.grad[0]).all() and torch.isfinite(logits.grad[1, 0]).all()
    y_ok = torch.tensor([[1, 2, 3, 99], [4, 7, 5, 6]], device=DEV)
    loss2, _ = masked_ce(logits.detach(), y_ok, torch.tensor([3, 4]), 1)
    assert torch.isfinite(loss2)
