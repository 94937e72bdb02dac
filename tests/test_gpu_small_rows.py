"""GPU tests of the single-step module APIs that external callers of the reference use (Rewriter-style drivers, src/lmtrain.py:221-240)
and of the device-side transcript cut -- the rows of SURVEY 8 that the whole-model goldens do not reach directly:
  * AutoRegDecoderLSTMCell.forward (src/modules.py:340-365) and the standalone LSTM-cell pointwise op, against the oracle's lstm_cell;
  * MultiheadCrossAttention.wrapup_encodings / forward(return_wgts, init_wgts_mask) (src/models.py:129-192): cached attributes,
    context, weights, gradients, against the oracle's attention_wrapup / attention_step;
  * las_b200.decode.greedy_transcripts against idx_to_str (src/infer.py:19-32)."""
import numpy as np
import pytest
import torch

from helpers import gu, orc, rel_err

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
TOL = 1e-4


def test_lstm_cell_pointwise_vs_formula():
    from las_b200.cellop import lstm_cell_pointwise
    g = torch.Generator().manual_seed(3)
    B, H = 5, 48
    gates = torch.randn(B, 4 * H, generator=g)
    c_prev = torch.randn(B, H, generator=g)
    mask = (torch.rand(B, H, generator=g) > 0.3).float() / 0.7
    wh, wc = torch.randn(B, H, generator=g), torch.randn(B, H, generator=g)
    go, co = gates.clone().requires_grad_(True), c_prev.clone().requires_grad_(True)
    i, f, gg, o = go[:, :H], go[:, H:2 * H], go[:, 2 * H:3 * H], go[:, 3 * H:]
    c_ref = torch.sigmoid(f) * co + torch.sigmoid(i) * torch.tanh(gg)
    h_ref = torch.sigmoid(o) * torch.tanh(c_ref) * mask
    ((h_ref * wh).sum() + (c_ref * wc).sum()).backward()
    gc, cc = gates.clone().to(DEV).requires_grad_(True), c_prev.clone().to(DEV).requires_grad_(True)
    h, c = lstm_cell_pointwise(gc, cc, mask.to(DEV))
    ((h * wh.to(DEV)).sum() + (c * wc.to(DEV)).sum()).backward()
    assert rel_err(h.detach().cpu().numpy(), h_ref.detach().numpy()) < TOL
    assert rel_err(c.detach().cpu().numpy(), c_ref.detach().numpy()) < TOL
    assert rel_err(gc.grad.cpu().numpy(), go.grad.numpy()) < TOL
    assert rel_err(cc.grad.cpu().numpy(), co.grad.numpy()) < TOL


@pytest.mark.parametrize('p_drop', [0.0, 0.3])
def test_autoreg_decoder_cell_forward_vs_oracle(p_drop):
    """Two stacked cells on cat[emb, ctx]; the DROPPED h of cell 0 is both cell 1's input and the stored state (src/modules.py:350-363)."""
    from las_b200.modules import AutoRegDecoderLSTMCell, set_mask_override
    P, E, DH, DO, B = 16, 32, 40, 24, 4
    rng = np.random.default_rng(9)
    cell = AutoRegDecoderLSTMCell(att_proj_dim=P, dec_emb_dim=E, dec_hid_dim=DH, dec_out_dim=DO, dec_mid_dropout=p_drop).to(DEV).train()
    sd = {k: torch.from_numpy(rng.uniform(-0.3, 0.3, size=tuple(v.shape)).astype(np.float32)) for k, v in cell.state_dict().items()}
    cell.load_state_dict(sd)
    emb = torch.from_numpy(rng.standard_normal((B, E)).astype(np.float32))
    ctx = torch.from_numpy(rng.standard_normal((B, P)).astype(np.float32))
    st = [torch.from_numpy(rng.standard_normal((B, n)).astype(np.float32) * 0.5) for n in (DH, DH, DO, DO)]
    m0 = torch.from_numpy(((rng.random((B, DH)) > p_drop) / (1 - p_drop)).astype(np.float32)) if p_drop else None
    m1 = torch.from_numpy(((rng.random((B, DO)) > p_drop) / (1 - p_drop)).astype(np.float32)) if p_drop else None
    wout = [torch.from_numpy(rng.standard_normal((B, n)).astype(np.float32)) for n in (DH, DH, DO, DO)]
    # oracle
    po = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    eo, co = emb.clone().requires_grad_(True), ctx.clone().requires_grad_(True)
    h0, c0 = orc.lstm_cell(torch.cat([eo, co], 1), st[0], st[1], po['lstms.0.weight_ih'], po['lstms.0.weight_hh'], po['lstms.0.bias_ih'],
                           po['lstms.0.bias_hh'])
    if m0 is not None:
        h0 = h0 * m0
    h1, c1 = orc.lstm_cell(h0, st[2], st[3], po['lstms.1.weight_ih'], po['lstms.1.weight_hh'], po['lstms.1.bias_ih'], po['lstms.1.bias_hh'])
    if m1 is not None:
        h1 = h1 * m1
    sum(( t * w).sum() for t, w in zip((h0, c0, h1, c1), wout)).backward()
    # module
    ed, cd = emb.clone().to(DEV).requires_grad_(True), ctx.clone().to(DEV).requires_grad_(True)
    if p_drop:
        set_mask_override(None, [m0, m1], None)
    try:
        out = cell(ed, cd, [(st[0].to(DEV), st[1].to(DEV)), (st[2].to(DEV), st[3].to(DEV))])
    finally:
        set_mask_override(None, None, None)
    (H0, C0), (H1, C1) = out
    sum((t * w.to(DEV)).sum() for t, w in zip((H0, C0, H1, C1), wout)).backward()
    for got, ref in ((H0, h0), (C0, c0), (H1, h1), (C1, c1)):
        assert rel_err(got.detach().cpu().numpy(), ref.detach().numpy()) < TOL
    assert rel_err(ed.grad.cpu().numpy(), eo.grad.numpy()) < TOL
    assert rel_err(cd.grad.cpu().numpy(), co.grad.numpy()) < TOL
    for k, prm in cell.named_parameters():
        assert rel_err(prm.grad.cpu().numpy(), po[k].grad.numpy()) < TOL, k


@pytest.mark.parametrize('heads,use_prior', [(1, False), (4, False), (2, True)])
def test_cross_attention_module_vs_oracle(heads, use_prior):
    from las_b200.models import MultiheadCrossAttention
    B, T, Denc, DO, P = 3, 29, 40, 24, 32
    lens = [29, 11, 20]
    rng = np.random.default_rng(heads)
    att = MultiheadCrossAttention(enc_out_dim=Denc, dec_out_dim=DO, proj_dim=P, heads=heads, dropout=0.0).to(DEV)
    sd = {k: torch.from_numpy(rng.uniform(-0.3, 0.3, size=tuple(v.shape)).astype(np.float32)) for k, v in att.state_dict().items()}
    att.load_state_dict(sd)
    enc = torch.from_numpy(rng.standard_normal((B, T, Denc)).astype(np.float32))
    dec_h = torch.from_numpy(rng.standard_normal((B, DO)).astype(np.float32))
    wctx = torch.from_numpy(rng.standard_normal((B, P)).astype(np.float32))
    prior = None
    if use_prior:
        prior = torch.zeros(B, heads, 1, T)
        prior[..., 5:17] = 1.0
    # oracle
    po = {'a.' + k: v.clone().requires_grad_(True) for k, v in sd.items()}
    eo, ho = enc.clone().requires_grad_(True), dec_h.clone().requires_grad_(True)
    keys, vals, pad = orc.attention_wrapup(po, eo, lens, heads, 'a.')
    ctx_o, w_o, q_o = orc.attention_step(po, ho, keys, vals, pad, heads, prior.squeeze(2) if use_prior else None, 'a.')
    (ctx_o * wctx).sum().backward()
    # module
    ed, hd = enc.clone().to(DEV).requires_grad_(True), dec_h.clone().to(DEV).requires_grad_(True)
    att.wrapup_encodings(ed, torch.tensor(lens))
    d = P // heads
    assert tuple(att.keys.shape) == (B, heads, d, T) and tuple(att.values.shape) == (B, heads, T, d)       # src/models.py:143-149
    assert tuple(att.masks.shape) == (B, heads, 1, T) and att.masks.dtype == torch.bool
    assert torch.equal(att.masks[:, 0, 0].cpu(), torch.arange(T).unsqueeze(0) >= torch.tensor(lens).unsqueeze(1))
    assert rel_err(att.keys.detach().permute(0, 1, 3, 2).cpu().numpy(), keys.detach().numpy()) < TOL
    assert rel_err(att.values.detach().cpu().numpy(), vals.detach().numpy()) < TOL
    if use_prior:
        ctx_c, w_c = att(hd, init_wgts_mask=prior.to(DEV))
        assert not w_c.requires_grad                       # the reference returns the detached pre-prior weights (:178,188)
    else:
        ctx_c, w_c = att(hd, return_wgts=True)
        assert tuple(att(hd).shape) == (B, P)              # return_wgts=False: context only
    assert tuple(att.queries.shape) == (B, heads, 1, d)
    assert tuple(w_c.shape) == (B, heads, 1, T)
    (ctx_c * wctx.to(DEV)).sum().backward()
    assert rel_err(ctx_c.detach().cpu().numpy(), ctx_o.detach().numpy()) < TOL
    assert np.abs(w_c.detach().squeeze(2).cpu().numpy() - w_o.detach().numpy()).max() < 1e-5
    assert rel_err(ed.grad.cpu().numpy(), eo.grad.numpy()) < TOL
    assert rel_err(hd.grad.cpu().numpy(), ho.grad.numpy()) < TOL
    gmax = max(float(v.grad.abs().max()) for v in po.values() if v.grad is not None)
    for k, prm in att.named_parameters():
        if po['a.' + k].grad is None:
            assert prm.grad is None, k                     # final_map is never used
            continue
        if k == 'key_map.bias':
            # mathematically zero (a constant added to every energy of a row cancels in the softmax): both sides hold round-off only
            assert float(prm.grad.abs().max()) < 1e-5 * gmax and float(po['a.' + k].grad.abs().max()) < 1e-5 * gmax
            continue
        assert rel_err(prm.grad.cpu().numpy(), po['a.' + k].grad.numpy(), 1e-3 * gmax) < TOL, k


def test_device_side_transcripts_match_idx_to_str():
    """las_b200.decode.greedy_transcripts == [idx_to_str(pl.argmax(-1), VOCAB, SOS, EOS) for pl in pred_logits] (src/infer.py:19-32,66):
    every <sos> skipped, cut at the first <eos>; rows with no <eos>, rows that start with <eos>, steps not a multiple of 32."""
    from las_b200.decode import greedy_transcripts, transcript_cut
    rng = np.random.default_rng(4)
    B, steps, V = 37, 77, 30
    chars = rng.integers(0, 29, size=(B, steps))                       # 0 = <sos> appears inside sequences
    for b in range(B):
        if b % 3 == 0:
            chars[b, rng.integers(0, steps)] = 29                       # an <eos> somewhere
        if b % 7 == 0:
            chars[b, rng.integers(0, steps, size=3)] = 29               # several: the first one counts
    chars[1, 0] = 29                                                    # empty transcript
    chars[2] = np.where(chars[2] == 29, 5, chars[2])                    # no <eos> at all
    chars[4, 31], chars[5, 32], chars[6, 33] = 29, 29, 29               # around the 32-step ballot boundary
    want = [orc.idx_to_str(chars[b], orc.VOCAB, 0, 29) for b in range(B)]
    got = greedy_transcripts(torch.from_numpy(chars.T.copy()).to(DEV), orc.VOCAB, 0, 29)          # (steps, B): Speller.last_chars layout
    assert got == want
    logits = torch.full((B, steps, V), -1.0)
    logits.scatter_(2, torch.from_numpy(chars).unsqueeze(-1), 1.0)
    assert greedy_transcripts(logits.to(DEV), orc.VOCAB, 0, 29) == want
    toks, lens = transcript_cut(torch.from_numpy(chars.T.copy()).to(DEV), 0, 29)
    assert lens.cpu().tolist() == [len(s) for s in want]


def test_transcripts_of_a_real_decode_equal_the_reference_host_loop():
    cfg = gu.get_config('tiny')
    from las_b200.models import ListenAttendSpell
    from las_b200.decode import greedy_transcripts
    m = ListenAttendSpell(**cfg).to(DEV)
    m.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in gu.make_state_dict(cfg, 707, scale=2.0).items()})
    m.eval()
    x, lx, _ = gu.make_inputs(708, 4, 400, 8, [400, 380, 333, 251])
    with torch.no_grad():
        logits, _ = m(torch.from_numpy(x).to(DEV), torch.from_numpy(lx))
    host_loop = [orc.idx_to_str(pl.argmax(-1).cpu().numpy(), orc.VOCAB, 0, 29) for pl in logits]
    assert greedy_transcripts(m.spell.last_chars, orc.VOCAB, 0, 29) == host_loop
