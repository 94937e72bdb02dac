"""Ragged-batch soak of the shipped schedule (scripts/soak_train.py): every step a different (B, T, L, lengths, tf_rate) through the
reducer / backward overlap / progress-counter tiles / persistent decoder kernel -- nothing hangs, the loss stays finite, the
allocator settles."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ragged_batches_soak():
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'scripts', 'soak_train.py'), '40'], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert 'soak ok: 40 ragged steps' in r.stdout
