"""Rewriter (reference src/lmtrain.py:95-253): oracle restatement and CUDA path against golden vectors
generated from the unmodified reference class (oracle/make_golden.py::rewriter_case)."""
import numpy as np
import pytest
import torch

from helpers import gu, orc, load_golden, rel_err

TOL = 1e-4


def _oracle_params(sd):
    p = {k: torch.from_numpy(v.copy()).requires_grad_(True) for k, v in sd.items() if k != 'cls.weight'}
    return p


def test_oracle_rewriter_train_matches_reference_golden():
    g = load_golden('rewriter_train')
    cfg = gu.get_rewriter_config(str(g['cfg_name']))
    sd = gu.make_rewriter_state_dict(cfg, int(g['seed']), float(g['scale']))
    p = _oracle_params(sd)
    y = torch.from_numpy(g['y'])
    logits, att = orc.rewriter_forward(p, torch.from_numpy(g['x']), g['lx'].tolist(), enc_layers=cfg['enc_lstm_layers'],
                                       heads=cfg['att_heads'], training=True, steps=y.shape[1])
    loss = torch.nn.functional.cross_entropy(logits.reshape(-1, logits.shape[-1]), y.reshape(-1))
    loss.backward()
    assert rel_err(logits.detach().numpy(), g['logits']) < TOL
    assert np.abs(att.numpy() - g['att']).max() < 1e-5
    assert abs(float(loss) - float(g['loss'])) < 1e-5
    assert int(g['n_coins']) == y.shape[1] - 1                 # one coin per step t > 0, drawn and never used
    nograd = set(str(s) for s in g['nograd'])
    gmax = max(float(np.abs(g[k]).max()) for k in g.files if k.startswith('grad.'))
    for k, v in p.items():
        if k in nograd:
            assert v.grad is None, k
            continue
        assert rel_err(v.grad.numpy(), g['grad.' + k], 1e-3 * gmax) < TOL, k


def test_oracle_rewriter_greedy_matches_reference_golden():
    g = load_golden('rewriter_greedy')
    cfg = gu.get_rewriter_config(str(g['cfg_name']))
    sd = gu.make_rewriter_state_dict(cfg, int(g['seed']), float(g['scale']))
    p = {k: torch.from_numpy(v.copy()) for k, v in sd.items()}
    with torch.no_grad():
        logits, att = orc.rewriter_forward(p, torch.from_numpy(g['x']), g['lx'].tolist(), enc_layers=cfg['enc_lstm_layers'],
                                           heads=cfg['att_heads'], training=False, steps=cfg['CHR_MAX_STEPS'])
    assert np.array_equal(logits.argmax(-1).numpy(), g['chars'])
    assert rel_err(logits.numpy(), g['logits']) < TOL


def test_rewriter_state_dict_contract():
    from las_b200.lm import Rewriter
    for name in ('rw_micro', 'rw_yml'):
        cfg = gu.get_rewriter_config(name)
        m = Rewriter(**cfg)
        want = dict(gu.rewriter_state_dict_shapes(cfg))
        want['cls.weight'] = want['char_emb.weight']
        sd = m.state_dict()
        assert set(sd.keys()) == set(want.keys())
        for k, v in sd.items():
            assert tuple(v.shape) == tuple(want[k]), k
        assert [k for k, _ in m.named_parameters()] == [k for k, _ in gu.rewriter_state_dict_shapes(cfg)]
        assert m.cls.weight is m.char_emb.weight
    import src.lmtrain as shim                                  # the drop-in module path of the reference
    assert shim.Rewriter is Rewriter


def _cuda_model(cfg, sd, train):
    from las_b200.lm import Rewriter
    m = Rewriter(**cfg).to('cuda:0')
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    return m.train() if train else m.eval()


@pytest.mark.gpu
def test_rewriter_train_step_matches_reference_golden():
    g = load_golden('rewriter_train')
    cfg = gu.get_rewriter_config(str(g['cfg_name']))
    sd = gu.make_rewriter_state_dict(cfg, int(g['seed']), float(g['scale']))
    m = _cuda_model(cfg, sd, True)
    y = torch.from_numpy(g['y']).to('cuda:0')
    torch.manual_seed(int(g['seed']))
    state0 = torch.get_rng_state()
    logits, att = m(torch.from_numpy(g['x']).to('cuda:0'), torch.from_numpy(g['lx']), y, 0.5)
    loss = torch.nn.functional.cross_entropy(logits.reshape(-1, logits.shape[-1]), y.reshape(-1))
    loss.backward()
    assert rel_err(logits.detach().cpu().numpy(), g['logits']) < TOL
    assert not att.is_cuda and tuple(att.shape) == g['att'].shape
    assert np.abs(att.numpy() - g['att']).max() < 1e-5
    assert abs(float(loss) - float(g['loss'])) < 1e-5
    # RNG parity: exactly L-1 host coins were consumed
    torch.set_rng_state(state0)
    for _ in range(y.shape[1] - 1):
        torch.rand(1)
    expect = torch.rand(1)
    torch.set_rng_state(state0)
    m.zero_grad(set_to_none=True)
    m(torch.from_numpy(g['x']).to('cuda:0'), torch.from_numpy(g['lx']), y, 0.5)
    assert torch.equal(torch.rand(1), expect)
    # gradients (second forward above did not run backward: grads were cleared, recompute)
    logits, _ = m(torch.from_numpy(g['x']).to('cuda:0'), torch.from_numpy(g['lx']), y, 0.5)
    torch.nn.functional.cross_entropy(logits.reshape(-1, logits.shape[-1]), y.reshape(-1)).backward()
    nograd = set(str(s) for s in g['nograd'])
    gmax = max(float(np.abs(g[k]).max()) for k in g.files if k.startswith('grad.'))
    for k, prm in m.named_parameters():
        if k in nograd:
            assert prm.grad is None, k
            continue
        assert rel_err(prm.grad.cpu().numpy(), g['grad.' + k], 1e-3 * gmax) < TOL, k


@pytest.mark.gpu
def test_rewriter_greedy_identical_to_reference():
    g = load_golden('rewriter_greedy')
    cfg = gu.get_rewriter_config(str(g['cfg_name']))
    sd = gu.make_rewriter_state_dict(cfg, int(g['seed']), float(g['scale']))
    m = _cuda_model(cfg, sd, False)
    with torch.no_grad():
        logits, att = m(torch.from_numpy(g['x']).to('cuda:0'), torch.from_numpy(g['lx']))
    assert np.array_equal(logits.argmax(-1).cpu().numpy(), g['chars'])
    assert np.array_equal(m.last_chars.t().cpu().numpy(), g['chars'])
    assert rel_err(logits.cpu().numpy(), g['logits']) < TOL
    assert np.abs(att.numpy() - g['att']).max() < 1e-5


@pytest.mark.gpu
def test_rewriter_yml_dims_bf16_vs_oracle():
    """config/rewriter.yml dims (emb 256, 2x BiLSTM 256, P 128 x 4 heads, dec 256/128) in AMP mode against the fp32 oracle."""
    cfg = gu.get_rewriter_config('rw_yml', enc_dropouts=[0.0, 0.0], dec_lstm_dropout=0.0)
    sd = gu.make_rewriter_state_dict(cfg, 77)
    B, Tx, L = 4, 24, 5
    x, lx, y = gu.make_token_inputs(78, B, Tx, L, [24, 17, 9, 20])
    m = _cuda_model(cfg, sd, True)
    yd = torch.from_numpy(y).to('cuda:0')
    with torch.autocast('cuda', dtype=torch.bfloat16):
        logits, _ = m(torch.from_numpy(x).to('cuda:0'), torch.from_numpy(lx), yd, 0.5)
    logits = logits.detach()
    p = {k: torch.from_numpy(v.copy()) for k, v in sd.items()}
    with torch.no_grad():
        ol, _ = orc.rewriter_forward(p, torch.from_numpy(x), lx.tolist(), enc_layers=2, heads=4, training=True, steps=L)
    # the first step's logits do not depend on argmax feedback; later steps may legitimately diverge after a near-tie
    assert np.abs(logits[:, 0].float().cpu().numpy() - ol[:, 0].numpy()).max() < 2e-2
    same = (logits.argmax(-1).cpu() == ol.argmax(-1)).all(dim=1)
    for b in range(B):
        if bool(same[b]):
            assert np.abs(logits[b].float().cpu().numpy() - ol[b].numpy()).max() < 3e-2
