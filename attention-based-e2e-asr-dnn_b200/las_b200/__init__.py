"""las_b200 -- B200-native (sm_100a) implementation of the Listen-Attend-Spell training / greedy-decode hot path of
Astromsoc/attention-based-e2e-asr-dnn, behind the reference's nn.Module API.

    from las_b200.models import ListenAttendSpell          # or: put this directory on sys.path and `import src.models`

Importing the package does not need a GPU; running a module does (there is no CPU fallback).
"""
from . import _lib
from .models import Listener, ListenAttendSpell, MultiheadCrossAttention, Speller
from .modules import AutoRegDecoderLSTMCell, LockedLSTM, pyramLockedLSTM, set_mask_override
from .optim import FusedAdamW

__all__ = ['Listener', 'ListenAttendSpell', 'MultiheadCrossAttention', 'Speller', 'AutoRegDecoderLSTMCell', 'LockedLSTM',
           'pyramLockedLSTM', 'FusedAdamW', 'set_mask_override', 'launch_count', 'reset_launch_count']


def launch_count() -> int:
    """Kernels launched by liblas_b200.so since the last reset."""
    return int(_lib.load().las_launch_count())


def reset_launch_count() -> None:
    _lib.load().las_launch_count_reset()
