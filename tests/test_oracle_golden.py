"""CPU tests: the oracle restatement (oracle/las_oracle.py) against the golden fixtures produced by the unmodified
reference (oracle/make_golden.py).  This is what pins the oracle."""
import os

import numpy as np
import pytest
import torch

from helpers import (gu, orc, load_golden, fixture_cfg, oracle_train_from_fixture, rel_err, grad_floor)

TRAIN_CASES = ['micro_train_tf1', 'micro_train_tf05', 'micro_train_dropout', 'micro_train_initforce', 'tiny_train_tf1']
TOL = 1e-4      # north_star: fp32 logits and gradients within 1e-4 relative error


@pytest.mark.parametrize('name', TRAIN_CASES)
def test_oracle_train_matches_reference(name):
    g = load_golden(name)
    logits, att, loss, grads = oracle_train_from_fixture(g, torch.float32)
    assert logits.shape == g['logits'].shape
    assert tuple(att.shape) == g['att'].shape
    assert rel_err(logits.numpy(), g['logits']) < TOL
    assert np.abs(att.numpy() - g['att']).max() < 1e-5
    assert abs(float(loss) - float(g['loss'])) < 1e-5
    nograd = set(str(s) for s in g['nograd'])
    assert nograd == {'spell.attention.final_map.weight', 'spell.attention.final_map.bias'}
    floor = grad_floor(g)
    for k, gr in grads.items():
        if k in nograd:
            assert gr is None
            continue
        ref_norm = float(g['gradnorm.' + k])
        got = float(np.linalg.norm(gr.numpy().astype(np.float64)))
        assert abs(got - ref_norm) <= TOL * max(ref_norm, floor), k
        if ('grad.' + k) in g.files:
            assert rel_err(gr.numpy(), g['grad.' + k], floor) < TOL, k


def test_oracle_float64_agrees_with_float32_reference():
    g = load_golden('micro_train_tf1')
    logits, _, _, _ = oracle_train_from_fixture(g, torch.float64)
    assert rel_err(logits.numpy(), g['logits']) < 1e-5


@pytest.mark.parametrize('name', ['micro_greedy', 'tiny_greedy'])
def test_oracle_greedy_transcripts_identical(name):
    g = load_golden(name)
    cfg = fixture_cfg(g)
    sd = gu.make_state_dict(cfg, int(g['seed']), scale=float(g['scale']))
    p = {k: torch.from_numpy(v.copy()) for k, v in sd.items()}
    lc, sc = cfg['listener_configs'], cfg['speller_configs']
    with torch.no_grad():
        logits, att = orc.las_forward(p, torch.from_numpy(g['x']), g['lx'].tolist(), lstm_layers=lc['lstm_layers'],
                                      plstm_layers=lc['plstm_layers'], heads=1, training=False, steps=sc['CHR_MAX_STEPS'])
    chars = logits.argmax(-1).numpy()
    assert np.array_equal(chars, g['chars'])            # greedy indices identical
    strs = [orc.idx_to_str(c, orc.VOCAB, 0, 29) for c in chars]
    assert strs == [str(s) for s in g['transcripts']]
    gold = [orc.idx_to_str(r, orc.VOCAB, 0, 29) for r in g['y']]
    assert [orc.levenshtein(a, b) for a, b in zip(strs, gold)] == g['ld'].tolist()
    assert rel_err(logits.numpy(), g['logits']) < TOL
    assert tuple(att.shape) == g['att'].shape


def test_oracle_optimizer_matches_torch():
    g = load_golden('optimizer_adamw_amsgrad')
    n, steps = int(g['n_params']), int(g['n_steps'])
    params = [torch.from_numpy(g[f'p0_{i}'].copy()) for i in range(n)]
    state = [dict() for _ in range(n)]
    for s in range(steps):
        scale = float(g[f'scale_{s}'])
        grads = [torch.from_numpy(g[f'g_{s}_{i}']) * scale if f'g_{s}_{i}' in g.files else None for i in range(n)]
        found_inf, _ = orc.optimizer_step(params, grads, state, lr=5e-4, weight_decay=5e-6, inv_scale=1.0 / scale)
        assert found_inf == (s == 3)
        for i in range(n):
            np.testing.assert_allclose(params[i].numpy(), g[f'p_{s}_{i}'], rtol=1e-6, atol=1e-7)
    for i in range(n - 1):
        assert state[i]['step'] == int(g[f'state_{i}_step'])
        np.testing.assert_allclose(state[i]['max_exp_avg_sq'].numpy(), g[f'state_{i}_max_exp_avg_sq'], rtol=1e-6, atol=1e-9)


def test_levenshtein_known_answers():
    assert orc.levenshtein('kitten', 'sitting') == 3
    assert orc.levenshtein('', 'abc') == 3
    assert orc.levenshtein('flaw', 'lawn') == 2
    assert orc.levenshtein('same', 'same') == 0


def test_nonpositive_length_raises_like_pack_padded_sequence():
    cfg = gu.get_config('micro')
    sd = {k: torch.from_numpy(v) for k, v in gu.make_state_dict(cfg, 1).items()}
    x = torch.zeros(2, 16, 15)
    with pytest.raises(RuntimeError):
        orc.listener_forward(sd, x, [16, 0], 1, 3)
    with pytest.raises(RuntimeError):      # 7 frames -> length 0 at the third pyramid level
        orc.listener_forward(sd, x, [16, 7], 1, 3)


def test_oracle_train_at_the_benchmarked_lengths():
    """BASELINE configs[1]'s lengths (T = 1600, L = 300), best config with the yml's dropouts (masks replayed): the oracle against
    the unmodified reference's logits, loss, gradient norms and strided gradient samples (about 20 s)."""
    g = load_golden('best_train_T1600_L300')
    torch.set_num_threads(os.cpu_count() or 8)
    logits, att, loss, grads = oracle_train_from_fixture(g, torch.float32)
    assert rel_err(logits.numpy(), g['logits']) < TOL
    assert np.abs(att.numpy() - g['att']).max() < 1e-5
    assert abs(float(loss) - float(g['loss'])) < 1e-5
    # the fixture's own conditioning: the reference in float64 (same masks) agrees with its fp32 run far inside the bar
    assert rel_err(g['logits'], g['logits64']) < 1e-5
    floor = grad_floor(g)
    for k, gr in grads.items():
        if gr is None:
            continue
        got = gr.numpy()
        ref_norm = float(g['gradnorm.' + k])
        assert abs(float(np.linalg.norm(got.astype(np.float64))) - ref_norm) <= TOL * max(ref_norm, floor), k
        stride = max(1, -(-got.size // 4096))
        assert np.abs(got.reshape(-1)[::stride] - g['gradsample.' + k]).max() <= TOL * max(float(g['gradabsmax.' + k]), floor), k


def test_oracle_greedy_at_the_benchmarked_lengths():
    """BASELINE configs[3]'s lengths (T = 3000 -> T_enc = 375, 600 greedy steps): transcripts identical to the reference."""
    g = load_golden('best_greedy_T3000')
    cfg = fixture_cfg(g)
    sd = gu.make_state_dict(cfg, int(g['seed']), scale=float(g['scale']))
    p = {k: torch.from_numpy(v.copy()) for k, v in sd.items()}
    torch.set_num_threads(os.cpu_count() or 8)
    with torch.no_grad():
        logits, att = orc.las_forward(p, torch.from_numpy(g['x']), g['lx'].tolist(), lstm_layers=1, plstm_layers=3, heads=1,
                                      training=False, steps=600)
    chars = logits.argmax(-1).numpy()
    assert np.array_equal(chars, g['chars'])
    assert [orc.idx_to_str(c, orc.VOCAB, 0, 29) for c in chars] == [str(s) for s in g['transcripts']]
    # max-norm error against the reference's float64 run, no worse than 3x the reference's own fp32 round-off (or the bar)
    own = rel_err(g['logits'], g['logits64'])
    assert rel_err(logits.numpy(), g['logits64']) < max(TOL, 3 * own)
    assert float(g['margin_min'].min()) > 10 * np.abs(g['logits'] - g['logits64']).max()      # transcripts are well separated
