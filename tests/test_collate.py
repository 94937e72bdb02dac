"""Loader collate + SpecAugment (reference src/utils.py:95-128): oracle restatement and the device-side
collator against golden vectors produced by the unmodified reference collate_fn (oracle/make_golden.py::collate_case)."""
import numpy as np
import pytest
import torch

from helpers import orc, load_golden

CASES = ['collate_specaug_short', 'collate_specaug_long', 'collate_plain']


def _batch(g):
    return [(torch.from_numpy(g[f'mfcc_{i}']), torch.from_numpy(g[f'trans_{i}'])) for i in range(int(g['n']))]


@pytest.mark.parametrize('name', CASES)
def test_oracle_collate_matches_reference_golden(name):
    g = load_golden(name)
    batch = _batch(g)
    torch.manual_seed(int(g['seed']))
    x, y, lx, ly = orc.collate_specaug([b[0] for b in batch], [b[1] for b in batch], bool(g['specaug']))
    assert np.array_equal(x.numpy(), g['x'])                 # bit-exact: padding, sort order and mask intervals
    assert np.array_equal(y.numpy(), g['y'])
    assert lx.tolist() == g['lx'].tolist() and ly.tolist() == g['ly'].tolist()


@pytest.mark.gpu
@pytest.mark.parametrize('name', CASES)
def test_device_collator_matches_reference_golden(name):
    from las_b200.data import DeviceCollator
    g = load_golden(name)
    coll = DeviceCollator('cuda:0', use_specaug=bool(g['specaug']))
    torch.manual_seed(int(g['seed']))
    x, y, lx, ly = coll(_batch(g))
    assert x.is_cuda and x.dtype == torch.float32
    assert np.array_equal(x.cpu().numpy(), g['x'])           # bit-exact
    assert np.array_equal(y.numpy(), g['y'])
    assert lx.tolist() == g['lx'].tolist() and ly.tolist() == g['ly'].tolist()
    # RNG parity: the collator consumed exactly the draws torchaudio would have
    torch.manual_seed(int(g['seed']))
    n_draws = 4 if bool(g['specaug']) else 0
    for _ in range(n_draws):
        torch.rand(1)
    expect = torch.rand(1)
    torch.manual_seed(int(g['seed']))
    coll(_batch(g))
    assert torch.equal(torch.rand(1), expect)
