"""GPU tests of the tcgen05 bf16 tensor-core GEMM (las_gemm_bf16_tc) and of the bf16 ("AMP") mode of the model.
Reference for the GEMM: torch fp32/fp64 matmul of the SAME bf16-rounded operands (so only accumulation order differs);
tolerance 2e-3 relative to the output scale.  For the model: north_star's AMP bar, logits within 2e-3 absolute of the fp32
reference."""
import numpy as np
import pytest
import torch

from helpers import gu, orc, load_golden, fixture_cfg, rel_err

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _rand_bf16(shape, seed):
    g = torch.Generator(device='cpu').manual_seed(seed)
    return torch.randn(*shape, generator=g).to(torch.bfloat16).to(DEV)


@pytest.mark.parametrize('B,R,K,N', [(1, 128, 64, 256), (1, 300, 192, 512), (3, 200, 128, 1024), (2, 77, 64, 100), (4, 129, 2048, 4096)])
def test_tc_gemm_forward_form(B, R, K, N):
    """C[b,r,:] = A[b,r,:] . W^T + bias1 + bias2 with A a strided (batch, row) view -- the forward gate projection."""
    from las_b200 import functional as LF
    Rpad = R + 3                                            # batch stride != R * row stride (padded batch)
    A = _rand_bf16((B, Rpad, K), 1)
    W = _rand_bf16((N, K), 2)
    b1 = torch.randn(N, device=DEV)
    b2 = torch.randn(N, device=DEV)
    C = torch.full((B, R, N), float('nan'), device=DEV)
    LF.gemm_tc(A, W, C, R, N, K, a_batches=B, a_s1=K, a_s2=Rpad * K, b_s1=K, c_bs=R * N, ldc=N, bias1=b1, bias2=b2)
    ref = A[:, :R].float().double() @ W.float().double().t() + (b1 + b2).double()
    assert torch.isfinite(C).all()
    assert rel_err(C.cpu().numpy(), ref.cpu().numpy()) < 2e-3


def test_tc_gemm_pyramid_concat_is_tma_geometry():
    """(B,T,D) -> drop odd frame -> (B,T//2,2D) as tensor-map strides (reference src/modules.py:171-185)."""
    from las_b200 import functional as LF
    B, T, D, N = 3, 271, 64, 256
    x = _rand_bf16((B, T, D), 3)
    W = _rand_bf16((N, 2 * D), 4)
    Tp = T // 2
    C = torch.empty(B, Tp, N, device=DEV)
    LF.gemm_tc(x, W, C, Tp, N, 2 * D, a_batches=B, a_s1=2 * D, a_s2=T * D, b_s1=2 * D, c_bs=Tp * N, ldc=N)
    ref = x[:, :2 * Tp].reshape(B, Tp, 2 * D).float() @ W.float().t()
    assert rel_err(C.cpu().numpy(), ref.cpu().numpy()) < 2e-3


def test_tc_gemm_length_tile_skip():
    from las_b200 import functional as LF
    B, R, K, N = 3, 400, 64, 256
    A = _rand_bf16((B, R, K), 5)
    W = _rand_bf16((N, K), 6)
    lens = torch.tensor([400, 130, 10], dtype=torch.int32, device=DEV)
    C = torch.zeros(B, R, N, device=DEV)
    LF.gemm_tc(A, W, C, R, N, K, a_batches=B, a_s1=K, a_s2=R * K, b_s1=K, c_bs=R * N, ldc=N, lens=lens)
    ref = A.float() @ W.float().t()
    for b, l in enumerate([400, 130, 10]):
        assert rel_err(C[b, :l].cpu().numpy(), ref[b, :l].cpu().numpy()) < 2e-3
    assert float(C[2, 128:].abs().max()) == 0.0           # tiles wholly past the length were never touched


@pytest.mark.parametrize('B,R,K,N', [(1, 128, 256, 256), (2, 130, 512, 64), (3, 64, 4096, 2048)])
def test_tc_gemm_dgrad_form(B, R, K, N):
    """dX[b,r,:] = dG[b,r,:] . W  with W (K, N) row-major as stored (MN-major B operand)."""
    from las_b200 import functional as LF
    dG = _rand_bf16((B, R, K), 7)
    W = _rand_bf16((K, N), 8)
    C = torch.empty(B, R, N, device=DEV)
    LF.gemm_tc(dG, W, C, R, N, K, a_batches=B, a_s1=K, a_s2=R * K, b_s1=N, b_mn=True, c_bs=R * N, ldc=N)
    ref = dG.float().double() @ W.float().double()
    assert rel_err(C.cpu().numpy(), ref.cpu().numpy()) < 2e-3


@pytest.mark.parametrize('B,T,M,N', [(1, 64, 128, 256), (3, 100, 256, 192), (2, 333, 4096, 2048), (5, 7, 128, 64)])
def test_tc_gemm_wgrad_form(B, T, M, N):
    """dW[m,n] = sum_{b,t} dG[b,t,m] X[b,t,n]: both operands MN-major, reduction over (batch, row) with OOB zero fill."""
    from las_b200 import functional as LF
    Tpad = T + 2
    dG = _rand_bf16((B, T, M), 9)
    X = _rand_bf16((B, Tpad, N), 10)
    C = torch.empty(M, N, device=DEV)
    # X is read with a +1 frame shift inside a padded batch, like the shifted h_{t-1} operand of dW_hh
    LF.gemm_tc(dG, X, C, M, N, T, k_batches=B, a_s1=M, a_s2=T * M, b_s1=N, b_s2=Tpad * N, ldc=N, a_mn=True, b_mn=True, b_off=N)
    ref = torch.einsum('btm,btn->mn', dG.float().double(), X[:, 1:T + 1].float().double())
    assert rel_err(C.cpu().numpy(), ref.cpu().numpy()) < 2e-3


def test_bf16_mode_lstm_layer_close_to_fp32_oracle():
    from las_b200 import functional as LF
    from las_b200.precision import set_precision
    rng = np.random.default_rng(5)
    H, D, B, T = 128, 64, 4, 50
    lens = [25, 20, 7, 25]
    x = torch.from_numpy(rng.standard_normal((B, T, D)).astype(np.float32))
    k = 1 / np.sqrt(H)
    names = ['weight_ih_l0', 'weight_hh_l0', 'bias_ih_l0', 'bias_hh_l0']
    shapes = [(4 * H, 2 * D), (4 * H, H), (4 * H,), (4 * H,)]
    p = {'l.' + n + suf: torch.from_numpy(rng.uniform(-k, k, size=s).astype(np.float32)) for suf in ['', '_reverse'] for n, s in zip(names, shapes)}
    po = {k_: v.clone().requires_grad_(True) for k_, v in p.items()}
    xo = x.clone().requires_grad_(True)
    yo = orc.bilstm_layer(xo.reshape(B, T // 2, 2 * D), lens, po, 'l.')
    wout = torch.from_numpy(rng.standard_normal(tuple(yo.shape)).astype(np.float32))
    (yo * wout).sum().backward()
    set_precision('bf16')
    try:
        pc = {k_: v.clone().to(DEV).requires_grad_(True) for k_, v in p.items()}
        xc = x.clone().to(DEV).requires_grad_(True)
        ws = [pc['l.' + n + suf] for suf in ['', '_reverse'] for n in names]
        yc = LF.lstm_layer(xc, torch.tensor(lens, dtype=torch.int32, device=DEV), max(lens), True, None, ws)
        (yc * wout.to(DEV)).sum().backward()
    finally:
        set_precision('auto')
    assert np.abs(yc.detach().cpu().numpy() - yo.detach().numpy()).max() < 2e-2
    assert rel_err(xc.grad.cpu().numpy(), xo.grad.numpy()) < 3e-2
    for k_ in p:
        assert rel_err(pc[k_].grad.cpu().numpy(), po[k_].grad.numpy()) < 3e-2, k_


@pytest.mark.parametrize('fuse_q', ['1', '0'])
def test_bf16_mode_best_config_dims_vs_oracle(fuse_q, monkeypatch):
    """AMP mode at the best config's dims (H 512, P 256, dec 512/256: the tcgen05 decoder GEMMs with split-K partials, the query
    projection fused into the cell-1 kernel, the DSMEM recurrence) against the fp32 oracle: logits within the AMP tolerance and
    the same gradients up to bf16 operand rounding.  fuse_q = 0 runs the separate query-projection GEMM instead."""
    from las_b200.models import ListenAttendSpell
    monkeypatch.setenv('LAS_DEC_FUSEQ', fuse_q)
    cfg = gu.get_config('best')
    sd = gu.make_state_dict(cfg, 11)
    B, T, L = 3, 72, 6
    x, lx, y = gu.make_inputs(12, B, T, L, [72, 49, 64])
    model = ListenAttendSpell(**gu.get_config('best')).to(DEV)
    model.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd.items()})
    model.train()
    yd = torch.from_numpy(y).to(DEV)
    with torch.autocast('cuda', dtype=torch.bfloat16):
        logits, _ = model(torch.from_numpy(x).to(DEV), torch.from_numpy(lx), yd, 1.0, False)
    loss = torch.nn.functional.cross_entropy(logits.float().reshape(-1, 30), yd.reshape(-1))
    loss.backward()
    p = {k: torch.from_numpy(v.copy()).requires_grad_(True) for k, v in sd.items() if k != 'spell.cls.weight'}
    p['spell.cls.weight'] = p['spell.char_emb.weight']
    ol, _ = orc.las_forward(p, torch.from_numpy(x), lx.tolist(), lstm_layers=1, plstm_layers=3, heads=1, training=True,
                            steps=L, dec_y=torch.from_numpy(y), coins=[True] * L)
    torch.nn.functional.cross_entropy(ol.reshape(-1, 30), torch.from_numpy(y).reshape(-1)).backward()
    err = np.abs(logits.detach().float().cpu().numpy() - ol.detach().numpy()).max()
    print(f'best dims, bf16 mode: max abs logit error vs fp32 oracle = {err:.3e}')
    assert err < 2e-3
    gmax = max(float(v.grad.abs().max()) for v in p.values() if v.grad is not None)
    for k, prm in model.named_parameters():
        if p[k].grad is None:
            continue
        assert rel_err(prm.grad.cpu().numpy(), p[k].grad.numpy(), 1e-2 * gmax) < 6e-2, k


@pytest.mark.parametrize('name', ['micro_train_tf1', 'tiny_train_tf1'])
def test_bf16_mode_logits_within_amp_tolerance(name):
    """north_star: AMP/bf16 logits within 2e-3 absolute of the reference."""
    from las_b200.models import ListenAttendSpell
    g = load_golden(name)
    cfg = fixture_cfg(g)
    sd = gu.make_state_dict(cfg, int(g['seed']))
    model = ListenAttendSpell(**cfg).to(DEV)
    model.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd.items()})
    model.train()
    y = torch.from_numpy(g['y']).to(DEV)
    with torch.autocast('cuda', dtype=torch.bfloat16):
        logits, _ = model(torch.from_numpy(g['x']).to(DEV), torch.from_numpy(g['lx']), y, 1.0, False)
    err = np.abs(logits.detach().float().cpu().numpy() - g['logits']).max()
    print(f'{name}: bf16-mode max abs logit error vs fp32 reference = {err:.3e}')
    assert err < 2e-3


@pytest.mark.parametrize('variant', ['default', 'tma'])
# the long sequences catch ordering races between a CTA's own epilogue and operands pushed by faster peers (seen once: T <= 21 passed)
@pytest.mark.parametrize('H,B,T,lens', [(64, 5, 9, [9, 3, 7, 1, 9]), (128, 40, 21, None), (512, 96, 12, None), (512, 130, 6, None),
                                        (128, 4, 400, None), (512, 96, 150, None), (128, 3, 1, None), (128, 3, 2, [2, 1, 2])])
def test_tc_recurrence_forward_vs_fp32_kernel(H, B, T, lens, variant, monkeypatch):
    """The tensor-pipe recurrence (bf16 operands) against the fp32 recurrence kernel on the same x-gates: same
    PackedSequence semantics (zeros past each length, reverse direction from each row's own end), values within bf16
    operand rounding."""
    import ctypes as C
    from las_b200 import _lib, functional as LF
    lib = _lib.load()
    # the two exchange schemes of the step-to-step h_t hand-off: SM-to-SM bulk copies inside a cluster (default when
    # the batch gives one chain per CTA) and the counter + TMA fallback
    monkeypatch.delenv('LAS_REC_DSMEM', raising=False)
    if variant != 'default':
        monkeypatch.setenv('LAS_REC_DSMEM', '0')
    rng = np.random.default_rng(H + B)
    if lens is None:
        lens = [T] + [int(v) for v in rng.integers(1, T + 1, size=B - 1)]
    ndir, F = 2, 2 * H
    assert lib.las_lstm_rec_tc_supported(B, H, ndir)
    gates0 = torch.from_numpy(rng.standard_normal((B, T, ndir, 4 * H)).astype(np.float32)).to(DEV)
    w_hh = torch.from_numpy((rng.uniform(-1, 1, size=(ndir, 4 * H, H)) / np.sqrt(H)).astype(np.float32)).to(DEV)
    lens_dev = torch.tensor(lens, dtype=torch.int32, device=DEV)
    mask = torch.from_numpy(((rng.random((B, F)) > 0.3) / 0.7).astype(np.float32)).to(DEV)
    res = []
    for tc in (False, True):
        gates = gates0.clone()
        hs = torch.full((B, T + 2, F), 7.0, device=DEV)
        cs = torch.full((B, T + 2, F), 7.0, device=DEV)
        out = torch.full((B, T, F), 7.0, device=DEV)
        if tc:
            wb = LF.cast_bf16(w_hh, ndir * 4 * H, H, H, H)
            nbytes = lib.las_lstm_rec_tc_workspace_bytes(B, H, ndir)
            ws = torch.empty(nbytes, dtype=torch.uint8, device=DEV)
            _lib.check(lib.las_lstm_rec_fwd_tc(gates.data_ptr(), wb.data_ptr(), lens_dev.data_ptr(), mask.data_ptr(), out.data_ptr(),
                                               hs.data_ptr(), cs.data_ptr(), B, T, H, ndir, 1, ws.data_ptr(), nbytes,
                                               torch.cuda.current_stream().cuda_stream), 'rec_fwd_tc')
        else:
            nbytes = lib.las_lstm_rec_workspace_bytes(B, H, ndir)
            ws = torch.empty(nbytes, dtype=torch.uint8, device=DEV)
            _lib.check(lib.las_lstm_rec_fwd_f32(gates.data_ptr(), w_hh.data_ptr(), lens_dev.data_ptr(), mask.data_ptr(), out.data_ptr(),
                                                hs.data_ptr(), cs.data_ptr(), B, T, H, ndir, ws.data_ptr(), nbytes,
                                                torch.cuda.current_stream().cuda_stream), 'rec_fwd')
        torch.cuda.synchronize()
        res.append((gates, hs, cs, out))
    (g0, h0, c0, o0), (g1, h1, c1, o1) = res
    assert np.abs((o1 - o0).cpu().numpy()).max() < 3e-2
    assert np.abs((h1 - h0).cpu().numpy()).max() < 3e-2
    assert np.abs((c1 - c0).cpu().numpy()).max() < 6e-2
    for b, l in enumerate(lens):
        if l < T:
            assert float(o1[b, l:].abs().max()) == 0.0 and float(h1[b, l + 1:].abs().max()) == 0.0
        # activated gates agree where the row is valid
        assert np.abs((g1[b, :l] - g0[b, :l]).cpu().numpy()).max() < 3e-2
    assert float(h1[:, 0].abs().max()) == 0.0 and float(h1[:, T + 1].abs().max()) == 0.0


@pytest.mark.parametrize('H,B,T,lens', [(64, 5, 9, [9, 3, 7, 1, 9]), (128, 40, 21, None), (512, 96, 12, None), (256, 130, 6, None),
                                        (128, 4, 300, None), (128, 3, 1, None), (128, 3, 2, [2, 1, 2])])
def test_tc_recurrence_backward_vs_fp32_kernel(H, B, T, lens):
    """BPTT on the tensor pipe against the fp32 BPTT kernel on identical saved activations."""
    from las_b200 import _lib, functional as LF
    lib = _lib.load()
    rng = np.random.default_rng(H + B + 1)
    if lens is None:
        lens = [T] + [int(v) for v in rng.integers(1, T + 1, size=B - 1)]
    ndir, F = 2, 2 * H
    st = torch.cuda.current_stream().cuda_stream
    gates0 = torch.from_numpy(rng.standard_normal((B, T, ndir, 4 * H)).astype(np.float32)).to(DEV)
    w_hh = torch.from_numpy((rng.uniform(-1, 1, size=(ndir, 4 * H, H)) / np.sqrt(H)).astype(np.float32)).to(DEV)
    lens_dev = torch.tensor(lens, dtype=torch.int32, device=DEV)
    mask = torch.from_numpy(((rng.random((B, F)) > 0.3) / 0.7).astype(np.float32)).to(DEV)
    dout = torch.from_numpy(rng.standard_normal((B, T, F)).astype(np.float32)).to(DEV)
    # forward once with the fp32 kernel to get consistent saved activations
    gates = gates0.clone()
    hs = torch.zeros(B, T + 2, F, device=DEV); cs = torch.zeros(B, T + 2, F, device=DEV); out = torch.zeros(B, T, F, device=DEV)
    nbytes = lib.las_lstm_rec_workspace_bytes(B, H, ndir)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=DEV)
    _lib.check(lib.las_lstm_rec_fwd_f32(gates.data_ptr(), w_hh.data_ptr(), lens_dev.data_ptr(), mask.data_ptr(), out.data_ptr(),
                                        hs.data_ptr(), cs.data_ptr(), B, T, H, ndir, ws.data_ptr(), nbytes, st), 'fwd')
    g_ref = gates.clone()
    _lib.check(lib.las_lstm_rec_bwd_f32(dout.data_ptr(), g_ref.data_ptr(), cs.data_ptr(), w_hh.data_ptr(), lens_dev.data_ptr(),
                                        mask.data_ptr(), B, T, H, ndir, ws.data_ptr(), nbytes, st), 'bwd_f32')
    g_tc = gates.clone()
    w_t = torch.empty(ndir, H, 4 * H, dtype=torch.bfloat16, device=DEV)
    _lib.check(lib.las_transpose_cast_bf16(w_hh.data_ptr(), w_t.data_ptr(), ndir, 4 * H, H, st), 'transpose')
    assert torch.equal(w_t.float(), w_hh.transpose(1, 2).to(torch.bfloat16).float())
    dgb = torch.full((B * T, ndir * 4 * H), float('nan'), dtype=torch.bfloat16, device=DEV)
    nb2 = lib.las_lstm_rec_tc_workspace_bytes(B, H, ndir)
    ws2 = torch.empty(nb2, dtype=torch.uint8, device=DEV)
    _lib.check(lib.las_lstm_rec_bwd_tc(dout.data_ptr(), g_tc.data_ptr(), dgb.data_ptr(), cs.data_ptr(), w_t.data_ptr(), lens_dev.data_ptr(),
                                       mask.data_ptr(), B, T, H, ndir, ws2.data_ptr(), nb2, st), 'bwd_tc')
    torch.cuda.synchronize()
    scale = float(g_ref.abs().max())
    assert np.abs((g_tc - g_ref).cpu().numpy()).max() < 2e-2 * scale
    assert torch.isfinite(dgb.float()).all()
    assert np.abs(dgb.float().view(B, T, ndir, 4 * H).cpu().numpy() - g_tc.cpu().numpy()).max() < 1e-2 * scale
    for b, l in enumerate(lens):
        if l < T:
            assert float(g_tc[b, l:].abs().max()) == 0.0
    # the form that accumulates the bias gradients inside the kernel (and skips the fp32 write-back): same bf16 dG, and the
    # per-slice rows add up to the column sums of the fp32 dG
    nsl = lib.las_lstm_rec_bwd_tc_dbias_slices(B, H, ndir)
    if nsl > 0:
        g_db = gates.clone()
        dgb2 = torch.full((B * T, ndir * 4 * H), float('nan'), dtype=torch.bfloat16, device=DEV)
        dbp = torch.full((ndir, nsl, 4 * H), float('nan'), device=DEV)
        _lib.check(lib.las_lstm_rec_bwd_tc_db(dout.data_ptr(), g_db.data_ptr(), dgb2.data_ptr(), cs.data_ptr(), w_t.data_ptr(),
                                              lens_dev.data_ptr(), mask.data_ptr(), B, T, H, ndir, ws2.data_ptr(), nb2, dbp.data_ptr(), st),
                   'bwd_tc_db')
        torch.cuda.synchronize()
        assert torch.equal(dgb2, dgb)
        db_ref = g_tc.double().sum(dim=(0, 1))                       # (ndir, 4H)
        db_got = dbp.double().sum(dim=1)
        assert float((db_got - db_ref).abs().max()) < 1e-4 * max(1.0, float(db_ref.abs().max()))


@pytest.mark.gpu
def test_backward_overlap_matches_plain_autograd():
    """Encoder weight gradients accumulated on the second stream beside the next layer's BPTT kernel (parameters whose gradients
    live in the reducer's buckets) == the same gradients returned through autograd (no reducer), bit for bit: same GEMMs, same
    split-K order, and 0 + g == g."""
    import copy
    from las_b200.models import ListenAttendSpell
    from las_b200.ddp import BucketedGradReducer
    from las_b200 import configs as gu
    from las_b200.loss import masked_ce
    cfg = gu.get_config('best')
    B, T, L = 8, 96, 12
    x, lx, y = gu.make_inputs(3, B, T, L)
    x, y, lx = torch.from_numpy(x).to(DEV), torch.from_numpy(y).to(DEV), torch.from_numpy(lx)
    ly = torch.full((B,), L, dtype=torch.int64)
    torch.manual_seed(5)
    m0 = ListenAttendSpell(**copy.deepcopy(cfg)).to(DEV).train()
    m1 = copy.deepcopy(m0)

    def run(model, use_reducer):
        red = BucketedGradReducer(list(model.named_parameters()), world_size=1) if use_reducer else None
        outs = []
        for it in range(2):                      # twice: the second pass reuses the pooled events / side stream
            if red is not None:
                red.zero_grad()
            else:
                model.zero_grad(set_to_none=True)
            torch.manual_seed(100 + it)
            with torch.autocast('cuda', dtype=torch.bfloat16):
                logits, _ = model(x, lx, y, 1.0, False)
            loss, _ = masked_ce(logits, y, ly)
            (loss * 1024.0).backward()
            if red is not None:
                red.finish()
            outs.append({n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None})
        torch.cuda.synchronize()
        return outs

    # the decoder's share of the overlap: the 13 parameters of the Speller loop (tied embedding / classifier weight, classifier bias,
    # both cells, query_map, init_query) get their gradients from las_speller_bwd_phases_f32(phases = 2) on the second stream beside the
    # top encoder layer's BPTT kernel, accumulated into the bucket views: same kernels on the same operands, and 0 + g == g
    speller_loop_params = ['spell.char_emb.weight', 'spell.cls.bias', 'spell.init_query', 'spell.attention.query_map.weight',
                           'spell.attention.query_map.bias'] + [f'spell.lstms.lstms.{i}.{w}_{k}' for i in (0, 1) for w in ('weight', 'bias')
                                                                for k in ('ih', 'hh')]
    names = dict(m0.named_parameters())
    assert all(n in names for n in speller_loop_params), [n for n in speller_loop_params if n not in names]
    from las_b200 import functional as LF
    calls = []
    lib = LF._lib.load()
    orig = lib.las_speller_bwd_phases_f32

    class _Spy:                                   # ctypes function pointers take no attributes: wrap the library object's entry
        def __call__(self, s_, g_, phases, stream):
            calls.append(int(phases))
            return orig(s_, g_, phases, stream)
    g_plain = run(m0, False)
    assert calls == []
    lib.las_speller_bwd_phases_f32 = _Spy()
    try:
        g_ovl = run(m1, True)
    finally:
        lib.las_speller_bwd_phases_f32 = orig
    assert calls == [1, 2, 1, 2], calls           # the deferred route was taken in both passes: loop first, parameter gradients later
    checked = 0
    for it in range(2):
        for n, g in g_plain[it].items():
            assert n in g_ovl[it], n
            if n.startswith('listen.') or n in speller_loop_params:
                assert torch.equal(g, g_ovl[it][n]), (it, n, (g - g_ovl[it][n]).abs().max().item())
                checked += 1
            else:
                torch.testing.assert_close(g, g_ovl[it][n], rtol=1e-5, atol=1e-6)
    assert checked >= 2 * (4 * 8 + len(speller_loop_params))


def test_tc_gemm_fp16_operands():
    """tcgen05 kind::f16 with IEEE fp16 operands (instruction-descriptor format bits 7..9 / 10..12 = 0): the decoder's forward
    activations and weights are fp16 (3 more mantissa bits than bf16, bounded range).  All three GEMM forms against fp64 matmul of
    the same rounded operands; a mixed fp16 x bf16 pair (an illegal instruction on B200, measured) is refused by the C ABI."""
    from las_b200 import functional as LF
    g = torch.Generator(device='cpu').manual_seed(31)
    B, R, K, N = 2, 150, 192, 320
    dt = torch.float16
    A = torch.randn(B, R, K, generator=g).to(dt).to(DEV)
    W = torch.randn(N, K, generator=g).to(dt).to(DEV)
    C = torch.empty(B, R, N, device=DEV)
    LF.gemm_tc(A, W, C, R, N, K, a_batches=B, a_s1=K, a_s2=R * K, b_s1=K, c_bs=R * N, ldc=N)
    ref = A.double() @ W.double().t()
    assert rel_err(C.cpu().numpy(), ref.cpu().numpy()) < 1e-5          # operands are exact: only fp32 accumulation differs
    Wk = torch.randn(K, N, generator=g).to(dt).to(DEV)                 # dgrad form (B MN-major)
    LF.gemm_tc(A, Wk, C, R, N, K, a_batches=B, a_s1=K, a_s2=R * K, b_s1=N, b_mn=True, c_bs=R * N, ldc=N)
    assert rel_err(C.cpu().numpy(), (A.double() @ Wk.double()).cpu().numpy()) < 1e-5
    X = torch.randn(B, R, N, generator=g).to(dt).to(DEV)               # wgrad form (both MN-major, reduction over (batch, row))
    Cw = torch.empty(K, N, device=DEV)
    LF.gemm_tc(A, X, Cw, K, N, R, k_batches=B, a_s1=K, a_s2=R * K, b_s1=N, b_s2=R * N, ldc=N, a_mn=True, b_mn=True)
    assert rel_err(Cw.cpu().numpy(), torch.einsum('brk,brn->kn', A.double(), X.double()).cpu().numpy()) < 1e-5
    with pytest.raises(RuntimeError, match='same 16-bit format'):
        LF.gemm_tc(A, W.to(torch.bfloat16), C, R, N, K, a_batches=B, a_s1=K, a_s2=R * K, b_s1=K, c_bs=R * N, ldc=N)


def _speller_vs_oracle(B, T, L, lx, tf_coins=None, seed=5):
    """best-config Speller in bf16 mode (the persistent decoder-step kernel) vs the fp32 oracle Speller, both on the encodings the
    fp32 Listener produces for a seeded batch (realistic operand magnitudes; the oracle does not have to run the encoder)."""
    from las_b200.models import ListenAttendSpell
    from las_b200.modules import set_mask_override
    cfg = gu.get_config('best')
    sd = gu.make_state_dict(cfg, seed)
    model = ListenAttendSpell(**gu.get_config('best')).to(DEV).train()
    model.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd.items()})
    x, lx, y = gu.make_inputs(seed + 1, B, T, L, lx=lx)
    with torch.no_grad():
        enc_h, enc_l = model.listen(torch.from_numpy(x).to(DEV), torch.from_numpy(lx))
    y = torch.from_numpy(y)
    tf = 1.0
    coins = [True] * L
    if tf_coins is not None:
        tf, coins = 0.5, [True] + list(tf_coins)
        set_mask_override(None, None, [0.1 if c else 0.9 for c in tf_coins])      # raw draws: <= 0.5 takes the gold token
    try:
        with torch.autocast('cuda', dtype=torch.bfloat16):
            logits, att = model.spell(enc_h, enc_l, y.to(DEV), tf, False)
    finally:
        set_mask_override(None, None, None)
    p = {k: torch.from_numpy(v.copy()) for k, v in sd.items()}
    with torch.no_grad():
        ol, oatt = orc.speller_forward(p, enc_h.cpu(), [int(v) for v in enc_l], heads=1, training=True, steps=L, dec_y=y, coins=coins)
    return logits.detach().float().cpu().numpy(), ol.numpy(), att.numpy(), oatt.numpy()


def test_persistent_decoder_more_rows_than_sms_and_two_slices_per_cell_cta():
    """B = 200 > 148 CTAs: attention CTAs own two batch rows, 7 batch slices over 6 slice groups -> cell CTAs own two slices (the greedy
    configuration's structure); ragged encoder lengths incl. 1."""
    from las_b200 import _lib
    B, T, L = 200, 320, 5
    assert _lib.load().las_speller_persistent(B, T // 8, 256, 512, 256, 30, 1, 0, 1) == 1
    rng = np.random.default_rng(3)
    lx = [T, 8] + [int(v) for v in rng.integers(9, T + 1, size=B - 2)]
    got, ref, att, oatt = _speller_vs_oracle(B, T, L, lx)
    assert np.abs(got - ref).max() < 2e-3
    assert np.abs(att - oatt).max() < 1e-3


def test_persistent_decoder_teacher_forcing_coins_and_argmax_feedback():
    """tf_rate 0.5 (the reference yml's default): per-step classifier + argmax inside the kernel, the coin decides per step whether
    the gold token or the fed-back argmax is embedded (src/models.py:354-358); device-side coin flags, no graph key."""
    B, T, L = 5, 240, 10
    coins = [True, False, False, True, False, True, False, False, True]          # steps 1..9
    got, ref, _, _ = _speller_vs_oracle(B, T, L, [240, 96, 200, 240, 56], tf_coins=coins)
    same = got.argmax(-1) == ref.argmax(-1)
    # identical fed-back tokens wherever the top-1 / top-2 gap is not a near-tie; logits within the AMP bar on agreeing prefixes
    for b in range(B):
        n = int(np.argmax(~same[b])) if (~same[b]).any() else L
        assert n >= L - 1 or np.sort(ref[b, n])[-1] - np.sort(ref[b, n])[-2] < 5e-3, (b, n)
        assert np.abs(got[b, :min(n + 1, L)] - ref[b, :min(n + 1, L)]).max() < 2e-3
    assert same.mean() > 0.9


def test_persistent_decoder_greedy_matches_launch_per_stage_loop(monkeypatch):
    """Eval mode (CHR_MAX_STEPS greedy steps, hist = 2 ring buffers): the persistent kernel (fp16 operands) against the
    launch-per-stage loop (LAS_DEC_PERSIST=0, bf16 operands) -- the fed-back argmax is the returned logits' argmax in both, first-step
    logits agree to operand rounding, transcripts agree except after a near-tie."""
    from las_b200.models import ListenAttendSpell
    cfg = gu.get_config('best', CHR_MAX_STEPS=25)
    sd = gu.make_state_dict(cfg, 9, scale=2.0)
    model = ListenAttendSpell(**gu.get_config('best', CHR_MAX_STEPS=25)).to(DEV).eval()
    model.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd.items()})
    B, T = 160, 264
    rng = np.random.default_rng(10)
    x, lx, _ = gu.make_inputs(11, B, T, 1, lx=[T] + [int(v) for v in rng.integers(8, T + 1, size=B - 1)])
    outs = []
    for persist in ('1', '0'):
        monkeypatch.setenv('LAS_DEC_PERSIST', persist)
        with torch.inference_mode():
            enc_h, enc_l = model.listen(torch.from_numpy(x).to(DEV), torch.from_numpy(lx))
            with torch.autocast('cuda', dtype=torch.bfloat16):
                lg, _ = model.spell(enc_h, enc_l)
        outs.append((lg.float().cpu().numpy(), model.spell.last_chars.t().cpu().numpy()))
    (la, ca), (lb, cb) = outs
    assert np.array_equal(ca, la.argmax(-1)) and np.array_equal(cb, lb.argmax(-1))
    scale = np.abs(lb[:, 0]).max()
    assert np.abs(la[:, 0] - lb[:, 0]).max() < 5e-3 * max(1.0, scale)
    assert (ca == cb).mean() > 0.9


@pytest.mark.parametrize('cfgname,T,lx', [('best', 72, [72, 40, 8]), ('best', 640, [640, 512, 77]), ('tiny', 96, [96, 61, 10])])
def test_fused_backward_step_matches_separate_launches(cfgname, T, lx, monkeypatch):
    """Backward decoder step: attention backward + dh1 = dq . Wq + LSTMCell-1 backward in ONE launch (csrc/attn_tail.h; one CTA per
    row at T_enc = 9, a CTA pair per row at T_enc = 80) against the three separate launches (LAS_BWD_FUSE_TAIL=0), same masks.  The fused
    form multiplies the fp32 dq by the bf16 weights, the separate GEMM rounds dq to bf16 first: not bit-identical, so both are also
    measured against the fp32-mode gradients of the same model and masks -- the fused step must not be further away."""
    from las_b200.models import ListenAttendSpell
    # 'tiny': P = 64, DH = 128, DO = 64 (one 64-column chunk, eight k groups in the fused dq . Wq) -- the small-dimension paths of both kernels
    sd = gu.make_state_dict(gu.get_config(cfgname), 21)
    B, L = len(lx), 7
    x, lxa, y = gu.make_inputs(22, B, T, L, lx=lx)
    ly = torch.tensor([L, L - 2, 3][:B])
    grads = {}
    for name, fuse, amp, att_tc in (('fp32', '1', False, '1'), ('fused', '1', True, '0'), ('separate', '0', True, '0'), ('fused_tc', '1', True, '1')):
        monkeypatch.setenv('LAS_BWD_FUSE_TAIL', fuse)
        monkeypatch.setenv('LAS_BWD_ATT_TC', att_tc)      # 1: fp16 K / V rows through mma.sync in the fused backward step (attn_bwd_tc_kernel)
        torch.manual_seed(5)
        model = ListenAttendSpell(**gu.get_config(cfgname, mid_dropout=0.3, dec_lstm_dropout=0.3)).to(DEV).train()
        model.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd.items()})
        with torch.autocast('cuda', dtype=torch.bfloat16, enabled=amp):
            logits, _ = model(torch.from_numpy(x).to(DEV), torch.from_numpy(lxa), torch.from_numpy(y).to(DEV), 1.0, False)
        mask = (torch.arange(L)[None, :] < ly[:, None]).to(DEV)
        loss = (torch.nn.functional.cross_entropy(logits.float().reshape(-1, 30), torch.from_numpy(y).to(DEV).reshape(-1), reduction='none')
                * mask.reshape(-1)).sum() / mask.sum()
        loss.backward()
        grads[name] = {k: p.grad.detach().cpu().numpy().copy() for k, p in model.named_parameters() if p.grad is not None}
    assert set(grads['fused']) == set(grads['separate']) == set(grads['fp32'])
    gmax = max(float(np.abs(v).max()) for v in grads['fp32'].values())
    err = {n: max((rel_err(grads[n][k], grads['fp32'][k], 1e-2 * gmax), k) for k in grads['fp32']) for n in ('fused', 'separate', 'fused_tc')}
    ab = max((rel_err(grads['fused'][k], grads['separate'][k], 1e-2 * gmax), k) for k in grads['fp32'])
    ac = max((rel_err(grads['fused_tc'][k], grads['fused'][k], 1e-2 * gmax), k) for k in grads['fp32'])
    print(f'backward step vs fp32 mode: fused {err["fused"]}, separate {err["separate"]}, tensor-pipe attention {err["fused_tc"]}; '
          f'fused vs separate {ab}; tensor-pipe vs FMA attention {ac}')
    assert ab[0] > 0.0 and ac[0] > 0.0, 'two runs took the same path'
    assert ab[0] < 1e-2 and ac[0] < 1e-2, (ab, ac)
    assert err['fused'][0] < 1.25 * err['separate'][0] + 1e-3, err
    assert err['fused_tc'][0] < 1.25 * err['separate'][0] + 1e-3, err


def test_forward_pipelining_is_bit_identical(monkeypatch):
    """The next layer's gate projection issued as time tiles on the second stream behind the recurrence's progress counters
    (functional.py "Forward pipelining") against the plain layer-after-layer order: same tiles of the same GEMM, so logits and every
    gradient must be bit-identical; ragged lengths, yml dropouts."""
    from las_b200 import functional as LF
    from las_b200.models import ListenAttendSpell
    sd = gu.make_state_dict(gu.get_config('best'), 31)
    B, T, L = 4, 1280, 5
    x, lxa, y = gu.make_inputs(32, B, T, L, lx=[1280, 1100, 640, 37])
    res = {}
    for pipe in ('1', '0', 'halves'):
        monkeypatch.setenv('LAS_FWD_PIPELINE', '0' if pipe == '0' else '1')
        monkeypatch.setenv('LAS_BWD_PIPELINE', '0' if pipe == '0' else '1')          # same scheme for each layer's dX GEMM beside its BPTT kernel
        # whole tiles behind BOTH sweeps are the same GEMM tiles as the unpipelined launch; the default (direction halves, the second one
        # accumulating) sums the same products in another order: checked below at the accuracy of the mode, and exactly at the level
        # of the gate pre-activations in test_direction_half_gate_tiles
        monkeypatch.setenv('LAS_FWD_KSPLIT', '1' if pipe == 'halves' else '0')
        LF.last_pipeline_stats.clear()
        torch.manual_seed(6)
        model = ListenAttendSpell(**gu.get_config('best', init_dropout=0.3, mid_dropout=0.3, final_dropout=0.35)).to(DEV).train()
        model.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd.items()})
        with torch.autocast('cuda', dtype=torch.bfloat16):
            logits, _ = model(torch.from_numpy(x).to(DEV), torch.from_numpy(lxa), torch.from_numpy(y).to(DEV), 1.0, False)
        logits.float().square().mean().backward()
        torch.cuda.synchronize()
        res[pipe] = (logits.detach().cpu().numpy().copy(), {k: p.grad.detach().cpu().numpy().copy() for k, p in model.named_parameters() if p.grad is not None},
                     dict(LF.last_pipeline_stats))
    stats = res['1'][2]
    print('pipelined tiles (early, late) per layer:', stats)
    assert sum(e for k, (e, _) in stats.items() if k[0] != 'bwd') >= 3, stats        # tiles that really ran beside a forward recurrence
    assert sum(e for k, (e, _) in stats.items() if k[0] == 'bwd') >= 3, stats        # ... and beside a BPTT kernel
    assert all(e == 0 for e, _ in res['0'][2].values()), res['0'][2]
    assert np.array_equal(res['1'][0], res['0'][0])
    for k in res['0'][1]:
        assert np.array_equal(res['1'][1][k], res['0'][1][k]), k
    # direction halves: more of them run beside the recurrences than whole tiles did; results agree at the accuracy of bf16 mode (a
    # different fp32 summation order flips a bf16 rounding of h here and there)
    hstats = res['halves'][2]
    print('direction halves (beside, after) per layer:', hstats)
    assert sum(e for k, (e, _) in hstats.items() if k[0] != 'bwd') > sum(e for k, (e, _) in stats.items() if k[0] != 'bwd'), (hstats, stats)
    dl = float(np.abs(res['halves'][0] - res['0'][0]).max())
    print('direction halves vs unpipelined: max abs logit difference', dl)
    assert dl < 2e-3, dl
    gmax = max(float(np.abs(v).max()) for v in res['0'][1].values())
    for k in res['0'][1]:
        ref = res['0'][1][k]
        # floor: key_map.bias has a mathematically zero gradient (a constant added to every energy), pure round-off on both sides
        assert np.abs(res['halves'][1][k] - ref).max() <= 5e-2 * np.abs(ref).max() + 1e-4 * gmax, k


def test_tc_gemm_chunked_reduction():
    """las_gemm_bf16_tc with k_chunk / k_chunk_stride: the reduction covers chunk c at column c * stride of BOTH operands -- one
    direction's half of a pyramid row [f(2t) b(2t) f(2t+1) b(2t+1)] against the matching columns of W_ih -- plain and accumulating."""
    from las_b200 import functional as LF
    g = torch.Generator(device='cpu').manual_seed(5)
    B, T, H, N = 3, 150, 128, 320                    # frames (B, 2T, 2H); rows = frame pairs, K = 4H
    x = torch.randn(B, 2 * T, 2 * H, generator=g).to(torch.bfloat16).to(DEV)
    W = torch.randn(N, 4 * H, generator=g).to(torch.bfloat16).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    rows = x.double().reshape(B, T, 4 * H)
    out = torch.empty(B, T, N, device=DEV)
    full = None
    for d in (1, 0):                                  # reverse half first (writes, + bias), forward half accumulates
        cols = torch.cat([torch.arange(d * H, (d + 1) * H), torch.arange(2 * H + d * H, 2 * H + (d + 1) * H)]).to(DEV)
        part = rows[:, :, cols] @ W.double()[:, cols].t()
        full = part + bias.double() if full is None else full + part
        LF.gemm_tc(x.view(B * 2 * T, 2 * H), W, out, T, N, 2 * H, a_batches=B, a_s1=4 * H, a_s2=2 * T * 2 * H, b_s1=4 * H, c_bs=T * N, ldc=N,
                   bias1=bias if d == 1 else None, accumulate=(d == 0), a_off=d * H, b_off=d * H, k_chunk=H, k_chunk_stride=2 * H)
        assert rel_err(out.cpu().numpy(), full.float().cpu().numpy()) < 1e-5
    ref = (rows @ W.double().t() + bias.double()).float()
    assert rel_err(out.cpu().numpy(), ref.cpu().numpy()) < 1e-5


def test_direction_half_gate_tiles(monkeypatch):
    """Gate pre-activations of the NEXT layer projected beside a BiLSTM layer's recurrence as direction halves
    (functional._issue_direction_halves) against whole tiles (LAS_FWD_KSPLIT=0) and against fp64 arithmetic on the layer's own bf16
    output: every (row, gate) exact up to the fp32 summation order; the layer output itself is bit-identical."""
    from las_b200 import functional as LF
    H, D, B, T = 128, 64, 40, 1100
    rng = np.random.default_rng(3)
    lens = [T, T - 1] + sorted([int(v) for v in rng.integers(5, T + 1, size=B - 2)], reverse=True)
    x = torch.from_numpy(rng.standard_normal((B, T, D)).astype(np.float32)).to(DEV)
    k = 1 / np.sqrt(H)
    shapes = [(4 * H, D), (4 * H, H), (4 * H,), (4 * H,)]
    ws = [torch.from_numpy(rng.uniform(-k, k, size=s_).astype(np.float32)).to(DEV) for _ in range(2) for s_ in shapes]
    shapes_n = [(4 * H, 4 * H), (4 * H, H), (4 * H,), (4 * H,)]              # the pyramid layer above: input = two frames x two directions
    wn = [torch.from_numpy(rng.uniform(-k, k, size=s_).astype(np.float32)).to(DEV) for _ in range(2) for s_ in shapes_n]
    lens_dev = torch.tensor(lens, dtype=torch.int32, device=DEV)
    Tn = T // 2
    got = {}
    for mode in ('1', '0'):
        monkeypatch.setenv('LAS_FWD_KSPLIT', mode)
        LF.last_pipeline_stats.clear()
        with torch.autocast('cuda', dtype=torch.bfloat16):
            y = LF.lstm_layer(x, lens_dev, T, False, None, ws, next_layer=dict(weights=wn, pyramid=True, T=Tn))
        pre = getattr(y, '_las_pregates', None)
        assert pre is not None, 'the next layer\'s gates were not projected beside the recurrence'
        pre.event.synchronize()
        torch.cuda.synchronize()
        got[mode] = (y.detach().clone(), pre.gates.detach().clone(), y._las_bf16[0].detach().clone(), dict(LF.last_pipeline_stats))
    print('direction halves (beside, after):', got['1'][3], ' whole tiles:', got['0'][3])
    assert sum(e for e, _ in got['1'][3].values()) > sum(e for e, _ in got['0'][3].values())
    assert torch.equal(got['1'][0], got['0'][0])
    y16 = got['1'][2].double().reshape(B, T, 2 * H)[:, :2 * Tn].reshape(B, Tn, 4 * H)
    wcat = torch.cat([wn[0], wn[4]]).to(torch.bfloat16).double()
    bsum = torch.cat([wn[2] + wn[3], wn[6] + wn[7]]).double()
    ref = (y16 @ wcat.t() + bsum).float().reshape(B, Tn, 2, 4 * H)
    scale = float(ref.abs().max())
    for mode in ('1', '0'):
        err = float((got[mode][1] - ref).abs().max())
        assert err <= 2e-5 * scale, (mode, err, scale)


def test_base_layer_weight_gradients_behind_their_own_bptt_kernel(monkeypatch):
    """Weight-gradient GEMMs of the long layers (more than two 256-step tiles; always the layer that ends the backward pass, which has no later
    BPTT kernel to hide them behind) run as time tiles (K-chunks accumulated in fp32) behind the progress counters of the layer's OWN BPTT
    kernel, direction by direction
    (functional._wgrad_tiles_behind_bptt).  Same products, different summation order than the one-GEMM form: equal to 1e-5 of the
    tensor's scale; everything else bit-identical."""
    import copy
    from las_b200 import functional as LF
    from las_b200.models import ListenAttendSpell
    from las_b200.ddp import BucketedGradReducer
    from las_b200 import configs as gcfg
    from las_b200.loss import masked_ce
    B, T, L = 4, 1280, 5
    x, lx, y = gcfg.make_inputs(41, B, T, L, lx=[1280, 1000, 700, 64])
    x, y, lx = torch.from_numpy(x).to(DEV), torch.from_numpy(y).to(DEV), torch.from_numpy(lx)
    ly = torch.full((B,), L, dtype=torch.int64)
    torch.manual_seed(8)
    m0 = ListenAttendSpell(**gcfg.get_config('best', mid_dropout=0.3)).to(DEV).train()
    res = {}
    for pipe in ('1', '0'):
        monkeypatch.setenv('LAS_BWD_WGRAD_PIPELINE', pipe)
        LF.last_pipeline_stats.clear()
        model = copy.deepcopy(m0)
        red = BucketedGradReducer(list(model.named_parameters()), world_size=1)
        red.zero_grad()
        torch.manual_seed(9)
        with torch.autocast('cuda', dtype=torch.bfloat16):
            logits, _ = model(x, lx, y, 1.0, False)
        loss, _ = masked_ce(logits, y, ly)
        (loss * 1024.0).backward()
        red.finish()
        torch.cuda.synchronize()
        res[pipe] = ({n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}, dict(LF.last_pipeline_stats))
    st = res['1'][1]
    print('tiles (early, late):', st)
    assert any(k[0] == 'wgrad' and e >= 4 for k, (e, _) in st.items()), st
    assert not any(k[0] == 'wgrad' for k in res['0'][1])
    for n, g in res['0'][0].items():
        h = res['1'][0][n]
        tiled = n.startswith('listen.base.')          # the layer that ends the backward pass (LAS_BWD_WGRAD_PIPELINE=2 would tile every long layer: measured slower)
        if tiled and 'bias' not in n:
            scale = g.abs().max().item()
            assert (g - h).abs().max().item() <= 1e-5 * scale + 1e-12, (n, (g - h).abs().max().item(), scale)
            assert not torch.equal(g, h) or scale == 0.0, n          # the tiled form really ran
        else:
            assert torch.equal(g, h), n
