"""CPU tests of the boundary: the C-ABI library loads and exports every symbol include/las_b200.h declares (no compute
calls without a GPU), the module API / state_dict contract matches the reference, and the product path fails loudly
without CUDA (no fallback)."""
import os
import re

import numpy as np
import pytest
import torch

from helpers import gu, ROOT


def test_library_exports_every_header_symbol():
    from las_b200 import _lib
    header = open(os.path.join(ROOT, 'include', 'las_b200.h')).read()
    header = re.sub(r'/\*.*?\*/', '', header, flags=re.S)
    declared = set(re.findall(r'\b(las_[a-z0-9_]+)\s*\(', header))
    assert declared, 'no declarations parsed'
    assert declared == set(_lib.SIGNATURES.keys()), declared ^ set(_lib.SIGNATURES.keys())
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.las_abi_version() == 1


def test_no_gpu_fails_loudly():
    from las_b200 import _lib
    lib = _lib.load()
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    assert lib.las_init(0) != 0
    assert b'no CPU fallback' in lib.las_last_error()


@pytest.mark.parametrize('name', ['micro', 'tiny', 'best'])
def test_state_dict_contract(name):
    from las_b200.models import ListenAttendSpell
    cfg = gu.get_config(name)
    model = ListenAttendSpell(**gu.get_config(name))
    sd = model.state_dict()
    want = dict(gu.state_dict_shapes(cfg))
    want['spell.cls.weight'] = want['spell.char_emb.weight']
    assert set(sd.keys()) == set(want.keys())
    for k, v in sd.items():
        assert tuple(v.shape) == tuple(want[k]), k
    # named_parameters order == the optimizer's param order
    assert [k for k, _ in model.named_parameters()] == [k for k, _ in gu.state_dict_shapes(cfg)]
    assert model.spell.cls.weight is model.spell.char_emb.weight
    # init_hiddens are NOT registered
    assert not any('init_hiddens' in k for k in sd)
    if name == 'best':
        assert sum(p.numel() for p in model.parameters()) == 37734686
    # a state_dict produced for the reference loads
    model.load_state_dict({k: torch.from_numpy(v) for k, v in gu.make_state_dict(cfg, 3).items()})


def test_public_attributes_touched_by_callers():
    from las_b200.models import ListenAttendSpell
    m = ListenAttendSpell(**gu.get_config('micro'))
    assert m.spell.dec_vocab_size == 30                                   # src/train.py:66
    for attr in ('init_dropout', 'mid_dropout', 'final_dropout'):         # src/train.py:467-469
        setattr(m.listen, attr, getattr(m.listen, attr) * 0.5)
    for attr in ('att_dropout', 'dec_emb_dropout', 'dec_lstm_dropout'):   # src/train.py:470-472
        setattr(m.spell, attr, getattr(m.spell, attr) * 0.5)
    m.spell.attention.dropout *= 0.5                                      # src/train.py:474
    assert m.speller_configs['enc_out_dim'] == 2 * m.listener_configs['uniform_hid_dim']   # src/models.py:512


def test_constructor_constraints():
    from las_b200.models import MultiheadCrossAttention, Speller
    with pytest.raises(AssertionError):
        MultiheadCrossAttention(proj_dim=10, heads=4)                     # src/models.py:87
    with pytest.raises(ValueError):
        Speller(att_proj_dim=16, dec_emb_dim=48, att_heads=1)


def test_cpu_tensors_raise_no_fallback():
    from las_b200.models import ListenAttendSpell
    m = ListenAttendSpell(**gu.get_config('micro'))
    x = torch.zeros(2, 16, 15)
    with pytest.raises(RuntimeError, match='CUDA'):
        m(x, torch.tensor([16, 16]), torch.ones(2, 3, dtype=torch.long), 1.0)


def test_length_contract_matches_pack_padded_sequence():
    from las_b200.modules import LockedLSTM
    m = LockedLSTM(15, 32)
    x = torch.zeros(2, 16, 15)
    with pytest.raises(RuntimeError, match='greater than 0'):
        m(x, torch.tensor([16, 0]))


def test_src_shim_is_drop_in():
    import src.models as sm
    import src.modules as smod
    import las_b200
    assert sm.ListenAttendSpell is las_b200.ListenAttendSpell
    assert smod.pyramLockedLSTM is las_b200.pyramLockedLSTM


@pytest.mark.reference
def test_shim_resolves_reference_drivers_and_our_models():
    """`python -m src.train` of the reference picks up OUR src.models / src.modules and ITS OWN src.utils / src.constants
    (authoring container only: needs /root/reference)."""
    import subprocess
    import sys
    if not os.path.isdir('/root/reference/src'):
        pytest.skip('reference checkout not present')
    code = ("import sys, types\n"
            "for n in ['torchsummaryX','Levenshtein','seaborn','matplotlib','matplotlib.pyplot','wandb']:\n"
            "    sys.modules.setdefault(n, types.ModuleType(n))\n"
            "sys.modules['torchsummaryX'].summary = lambda *a, **k: None\n"
            "sys.modules['Levenshtein'].distance = lambda a, b: 0\n"
            "sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']\n"
            "import src.models, src.constants, src.utils\n"
            "import las_b200\n"
            "assert src.models.ListenAttendSpell is las_b200.ListenAttendSpell\n"
            "assert src.constants.EOS_IDX == 29 and '/root/reference' in src.utils.__file__\n"
            "print('ok')\n")
    env = dict(os.environ, LAS_REFERENCE_SRC='/root/reference/src', PYTHONPATH=os.path.join(ROOT, 'attention-based-e2e-asr-dnn_b200'))
    out = subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True, cwd='/tmp')
    assert out.returncode == 0 and 'ok' in out.stdout, out.stderr[-2000:]


def test_batched_coin_draw_equals_sequential_draws():
    """Speller.forward draws the teacher-forcing coins with one torch.rand(n) call; the reference draws torch.rand(1) n times
    (src/models.py:357).  Same values, same generator state afterwards."""
    for n in (1, 7, 16, 299, 600):
        torch.manual_seed(11785)
        seq = torch.stack([torch.rand(1) for _ in range(n)]).flatten()
        after_seq = torch.rand(4)
        torch.manual_seed(11785)
        bat = torch.rand(n)
        after_bat = torch.rand(4)
        assert torch.equal(seq, bat) and torch.equal(after_seq, after_bat), n


def test_lazy_host_tensor_waits_once_then_acts_like_a_plain_tensor():
    """The attention map comes back as a CPU tensor whose copy may still be in flight (las_b200.functional.LazyHostTensor):
    the first use must wait for the copy's event, and everything a trainer does with the map must keep working."""
    import io
    import numpy as np
    import torch
    from las_b200.functional import LazyHostTensor

    class Ev:
        n = 0
        def synchronize(self):
            Ev.n += 1

    base = torch.arange(24, dtype=torch.float32).view(2, 3, 4)
    t = LazyHostTensor(base.clone(), Ev())
    assert isinstance(t, torch.Tensor) and Ev.n == 0
    assert t.device.type == 'cpu' and Ev.n == 1                   # any access waits ...
    assert tuple(t.shape) == (2, 3, 4) and Ev.n == 1              # ... once
    a = np.asarray(t)
    assert a.shape == (2, 3, 4) and np.array_equal(a, base.numpy())
    r = t[0] * 2 + 1
    assert type(r) is torch.Tensor and torch.equal(r, base[0] * 2 + 1)
    assert torch.equal(torch.stack([t, t]).sum(0), 2 * base)
    assert type(t.permute(1, 2, 0)) is torch.Tensor
    assert 'tensor' in repr(t)
    buf = io.BytesIO(); torch.save(t.clone(), buf); buf.seek(0)
    assert torch.equal(torch.load(buf), base)
    for fresh in (LazyHostTensor(base.clone(), Ev()),):
        n0 = Ev.n
        assert fresh.numpy().sum() == base.sum() and Ev.n == n0 + 1


def test_backward_overlap_eligibility_does_not_depend_on_an_attached_profiler(monkeypatch):
    """Round 1 switched the backward overlap off when a CUDA injection library (ncu / nsys) was in the environment, because a whole
    bench under ncu had not finished in time.  A one-step launch list with the overlap on completes (profiles/launches_r2_step_summary.csv),
    so the schedule is the same with and without a profiler; LAS_BWD_OVERLAP=0 is the only switch."""
    import las_b200.functional as LF
    import torch
    assert not hasattr(LF, '_profiler_attached')
    w = torch.nn.Parameter(torch.zeros(2)); w.grad = torch.zeros(2); w._las_bucketed = True
    monkeypatch.delenv('LAS_BWD_OVERLAP', raising=False)
    monkeypatch.setenv('CUDA_INJECTION64_PATH', '/opt/nvidia/nsight-compute/target/libcuda-injection64.so')
    with torch.no_grad():
        assert LF._overlap_ok((w,)) is True
        monkeypatch.setenv('LAS_BWD_OVERLAP', '0')
        assert LF._overlap_ok((w,)) is False
    monkeypatch.delenv('LAS_BWD_OVERLAP', raising=False)
    assert LF._overlap_ok((w,)) is False                    # grad mode on: an autograd.grad / double-backward caller takes the plain route


def test_bf16_shadow_is_dropped_when_the_tensor_is_written_to(monkeypatch):
    """lstm_layer leaves the kernel-written bf16 copy of its output on the tensor (functional._bf16_shadow); the next layer may
    only use it while the fp32 tensor has not been modified since, has the same shape, and the tensor-pipe mode is on."""
    import torch
    import las_b200.functional as LF
    monkeypatch.setattr(LF, 'use_tensor_cores', lambda: True)
    y = torch.zeros(2, 3, 4)
    assert LF._bf16_shadow(y) is None                       # nothing attached
    y16 = torch.zeros(2, 3, 4, dtype=torch.bfloat16)
    y._las_bf16 = (y16, y._version)
    assert LF._bf16_shadow(y) is y16
    assert LF._bf16_shadow(y[:, :2]) is None                # a view is another tensor object: no attribute
    y.mul_(2.0)                                             # in-place edit by the caller -> stale copy must not be used
    assert LF._bf16_shadow(y) is None
    z = torch.zeros(2, 3, 4)
    z._las_bf16 = (torch.zeros(2, 3, 8, dtype=torch.bfloat16), z._version)
    assert LF._bf16_shadow(z) is None                       # shape mismatch
    z._las_bf16 = (y16, z._version)
    monkeypatch.setattr(LF, 'use_tensor_cores', lambda: False)
    assert LF._bf16_shadow(z) is None                       # fp32 parity mode never uses it
    # inference tensors have no version counter (the reference's dev evaluation runs under torch.inference_mode()): nothing is
    # attached, nothing is looked up, nothing raises
    with torch.inference_mode():
        yi = torch.zeros(2, 3, 4)
        LF._attach_bf16_shadow(yi, torch.zeros(2, 3, 4, dtype=torch.bfloat16))
        assert not hasattr(yi, '_las_bf16')
        monkeypatch.setattr(LF, 'use_tensor_cores', lambda: True)
        assert LF._bf16_shadow(yi) is None
    yn = torch.zeros(2, 3, 4)
    LF._attach_bf16_shadow(yn, y16)
    assert LF._bf16_shadow(yn) is y16
