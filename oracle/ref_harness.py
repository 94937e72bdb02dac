"""
ORACLE -- TEST INFRASTRUCTURE ONLY.

Imports and drives the UNMODIFIED reference (staged by oracle/build_ref.sh into the git-ignored oracle/_ref/, or read from
/root/reference in the authoring container).  Shared by oracle/make_golden.py (fixtures), oracle/ref_runner.py (the
reference arm of bench.py, the GPU comparator, the AMP checks, the drop-in trainer / inference runs).

The reference's modules live in a top-level package called `src` -- the same name as the drop-in shim in
attention-based-e2e-asr-dnn_b200/src -- so a process that drives the reference must NOT have the product package on sys.path
(oracle/ref_runner.py is always started as its own process for that reason).
"""
from __future__ import annotations

import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
STUBS = os.path.join(HERE, 'ref_stubs')


def reference_root() -> str:
    """Directory that holds the reference's `src/` and `config/`: $LAS_REFERENCE, oracle/_ref (staged copy), /root/reference."""
    for cand in (os.environ.get('LAS_REFERENCE'), os.path.join(HERE, '_ref'), '/root/reference'):
        if cand and os.path.isdir(os.path.join(cand, 'src')):
            return cand
    raise FileNotFoundError('the reference is not staged: run `sh oracle/build_ref.sh` where /root/reference exists')


def reference_available() -> bool:
    try:
        reference_root()
        return True
    except FileNotFoundError:
        return False


def import_reference():
    """`import src.models` of the unmodified reference.  torchsummaryX / Levenshtein / seaborn / matplotlib are imported by
    the reference for non-numerical purposes and are absent from this image: oracle/ref_stubs/."""
    for p in (STUBS, reference_root()):
        if p not in sys.path:
            sys.path.insert(0, p)
    import src.models as ref_models
    assert 'las_b200' not in (getattr(ref_models, '__file__', '') or ''), 'the drop-in shim shadowed the reference'
    return ref_models


class Recorder:
    """Records, in call order, every tf coin (torch.rand(1)), locked-dropout mask (Tensor.bernoulli_ followed by
    an in-place div_) and nn.Dropout mask (F.dropout) the reference draws."""

    def __init__(self):
        self.coins, self.locked, self.drops = [], [], []

    def __enter__(self):
        import torch.nn.functional as F
        self._rand, self._bern, self._drop = torch.rand, torch.Tensor.bernoulli_, F.dropout
        rec = self

        def rand(*a, **k):
            r = rec._rand(*a, **k)
            if tuple(r.shape) == (1,):
                rec.coins.append(float(r.item()))
            return r

        def bern(self_, *a, **k):
            r = rec._bern(self_, *a, **k)
            rec.locked.append(r)            # later div_'ed in place -> holds the final mask
            return r

        def drop(x, p=0.5, training=True, inplace=False):
            if not training or p == 0.0:
                return x
            m = rec._drop(torch.ones_like(x), p, True, False)
            rec.drops.append(m)
            return x * m

        torch.rand, torch.Tensor.bernoulli_, F.dropout = rand, bern, drop
        return self

    def __exit__(self, *exc):
        import torch.nn.functional as F
        torch.rand, torch.Tensor.bernoulli_, F.dropout = self._rand, self._bern, self._drop


class Replayer:
    """Feeds recorded coins / masks back, in the same call order, to another run of the reference (a float64 re-run that
    measures the reference's own fp32 round-off, or a GPU / autocast run on a committed fixture).
    locked: final locked-dropout masks (0 or 1/keep); drops: nn.Dropout masks (0 or 1/keep); coins: raw torch.rand(1) draws."""

    def __init__(self, coins, locked, drops):
        self.coins, self.locked, self.drops = list(coins), [m.clone() for m in locked], list(drops)

    @classmethod
    def from_recorder(cls, rec):
        return cls(rec.coins, rec.locked, rec.drops)

    def __enter__(self):
        import torch.nn.functional as F
        self._rand, self._bern, self._drop = torch.rand, torch.Tensor.bernoulli_, F.dropout
        rep = self

        def rand(*a, **k):
            if a == (1,):
                return torch.tensor([rep.coins.pop(0)], dtype=torch.float64)
            return rep._rand(*a, **k)

        def bern(self_, *a, **k):
            # the recorded tensor is the FINAL mask (0 or 1/keep after the in-place div_): put the 0/1 pattern back
            m = rep.locked.pop(0)
            return self_.copy_((m != 0).to(device=self_.device, dtype=self_.dtype))

        def drop(x, p=0.5, training=True, inplace=False):
            if not training or p == 0.0:
                return x
            return x * rep.drops.pop(0).to(device=x.device, dtype=x.dtype)

        torch.rand, torch.Tensor.bernoulli_, F.dropout = rand, bern, drop
        return self

    def __exit__(self, *exc):
        import torch.nn.functional as F
        torch.rand, torch.Tensor.bernoulli_, F.dropout = self._rand, self._bern, self._drop


def to_double(model):
    model = model.double()
    # init_hiddens is a plain Python list of Parameters (src/models.py:275-281): nn.Module.double() does not see it
    model.spell.init_hiddens = [tuple(t.double() for t in h) for h in model.spell.init_hiddens]
    return model


def to_device(model, device):
    model = model.to(device)
    # same for .to(device): the reference's own trainer relies on `.to(device)` inside Speller.forward (src/models.py:338-343)
    return model
