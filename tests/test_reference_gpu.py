"""GPU tests against the UNMODIFIED reference running on the same GPU (oracle/_ref, staged by oracle/build_ref.sh; it travels with the
repository snapshot).  Every reference run is its own process (oracle/ref_runner.py): its top-level package is called `src`, like
the drop-in shim.  Skipped when the staged copy is absent.

  * the AMP contract: our bf16 mode and the reference under torch.autocast, both measured against the reference's fp32
    CPU fixture at the benchmarked lengths -- ours must be inside the 2e-3 bar AND at least as close as the reference's own autocast run;
  * greedy decoding in the mode bench.py times: per-utterance agreement with the fp32 transcripts, ours vs the reference's autocast run;
  * the drop-in run: the reference's own Trainer.train_epoch / evaluate_epoch / infer_one_checkpoint (src/train.py:105-258,
    src/infer.py:36-81) driving the B200-native modules through the `src` shim."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from helpers import ROOT, load_golden
import test_gpu_bench_shapes as shapes

pytestmark = pytest.mark.gpu
STAGED = os.path.isdir(os.path.join(ROOT, 'oracle', '_ref', 'src')) or os.path.isdir('/root/reference/src')
needs_ref = pytest.mark.skipif(not STAGED, reason='oracle/_ref not staged (sh oracle/build_ref.sh)')


def ref_runner(*args, timeout=900):
    env = {k: v for k, v in os.environ.items() if k != 'PYTHONPATH'}
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'oracle', 'ref_runner.py')] + [str(a) for a in args], env=env,
                         capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    lines = [l for l in out.stdout.splitlines() if l.startswith('{')]
    assert out.returncode == 0 and lines, out.stderr[-3000:]
    return json.loads(lines[-1])


@needs_ref
def test_amp_logits_vs_reference_under_autocast_at_benchmarked_lengths(tmp_path):
    g = load_golden('best_train_T1600_L300')
    errs = {}
    for amp in ('bf16', 'fp16'):
        out = tmp_path / f'ref_{amp}.npz'
        d = ref_runner('fixture', '--name', 'best_train_T1600_L300', '--device', 'cuda', '--amp', amp, '--out', out)
        errs[amp] = float(np.abs(np.load(out)['logits'] - g['logits']).max())
    d32 = ref_runner('fixture', '--name', 'best_train_T1600_L300', '--device', 'cuda', '--amp', 'none', '--out', tmp_path / 'ref_fp32.npz')
    model, logits, att, loss = shapes._run_train_fixture(g, amp=True)
    ours = float(np.abs(logits - g['logits']).max())
    shapes._record(test='amp_vs_reference_autocast', fixture='best_train_T1600_L300', ours_bf16_mode=ours, reference_bf16_autocast=errs['bf16'],
                   reference_fp16_autocast=errs['fp16'], reference_fp32_cudnn=d32['logits_abs_vs_cpu_fp32_reference'])
    assert d32['logits_abs_vs_cpu_fp32_reference'] < 1e-4          # the reference's GPU fp32 run reproduces its CPU fixture
    assert ours < shapes.AMP_TOL
    assert ours <= errs['bf16']                                     # closer to fp32 than the reference's own bf16 autocast run


@needs_ref
def test_greedy_agreement_vs_reference_under_autocast(tmp_path):
    g = load_golden('best_greedy_T3000')
    out = tmp_path / 'ref_greedy_bf16.npz'
    ref_runner('fixture', '--name', 'best_greedy_T3000', '--device', 'cuda', '--amp', 'bf16', '--out', out)
    ref_chars = np.load(out)['chars']
    ref_agree = [float((ref_chars[b] == g['chars'][b]).mean()) for b in range(ref_chars.shape[0])]
    model, logits, att = shapes._run_greedy_fixture(g, amp=True)
    chars = logits.argmax(-1)
    agree = [float((chars[b] == g['chars'][b]).mean()) for b in range(chars.shape[0])]
    shapes._record(test='greedy_agreement_vs_reference_autocast', ours_bf16_mode=agree, reference_bf16_autocast=ref_agree)
    # argmax feedback: one flipped near-tie changes the rest of an utterance.  Ours must agree with the fp32 transcripts at least as
    # well (on average) as the reference's own autocast run does
    assert float(np.mean(agree)) >= float(np.mean(ref_agree)) - 0.05, (agree, ref_agree)


@needs_ref
def test_reference_trainer_and_inference_run_on_the_shim():
    """Two batches of Trainer.train_epoch (tf_rate 0.5, yml dropouts, GradScaler, clip, AdamW-amsgrad), one evaluate_epoch (greedy +
    Levenshtein) and infer_one_checkpoint (state_dict round trip, transcripts, csv) of the UNMODIFIED reference drivers, once on the
    reference's own modules and once on las_b200 through the path shim of INTEGRATION.md."""
    common = ('trainer', '--device', 'cuda', '--amp', 'fp16', '--config', 'tiny', '--B', 6, '--T', 240, '--L', 12, '--max-steps', 24)
    ref = ref_runner(*common, '--shim', 0)
    ours = ref_runner(*common, '--shim', 1)
    shapes._record(test='trainer_drop_in', reference=ref, shim=ours)
    assert 'las_b200' not in ref['models_file'] and '/attention-based-e2e-asr-dnn_b200/src/' in ours['models_file'].replace(os.sep, '/')
    assert ours['train_file'] == ref['train_file']                 # the same (reference) driver in both runs
    assert ref['finite'] and ours['finite']
    assert ours['att_shape'] == ref['att_shape'] and ours['n_preds'] == ref['n_preds'] == 6
    # same init (seeded), different dropout / coin draws on the device: the first-epoch losses agree to a few per cent
    assert abs(ours['trn_loss'] - ref['trn_loss']) < 0.1 * ref['trn_loss']
    assert abs(ours['dev_loss'] - ref['dev_loss']) < 0.1 * ref['dev_loss']
