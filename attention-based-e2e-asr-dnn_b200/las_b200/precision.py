"""Arithmetic mode of the hot path.

  'fp32' : parity mode -- every GEMM is the fp32 FFMA kernel; matches the reference's fp32 path to 1e-4 relative.
  'bf16' : tensor-pipe mode -- the LSTM gate GEMMs (forward, dgrad, wgrad) run as tcgen05 bf16 tiles with fp32
           accumulation; recurrence state, gates and all reductions stay fp32.  This is the AMP contract
           (north_star: logits within 2e-3 absolute of the reference under autocast).
  'auto' : (default) 'bf16' inside a torch.autocast region -- where the reference's train.py / infer.py run the model
           (src/train.py:130, src/infer.py:59) -- and 'fp32' outside.
"""
from __future__ import annotations

import os

import torch

_MODE = os.environ.get('LAS_PRECISION', 'auto')


def set_precision(mode: str) -> None:
    global _MODE
    if mode not in ('fp32', 'bf16', 'auto'):
        raise ValueError(mode)
    _MODE = mode


def get_precision() -> str:
    return _MODE


def use_tensor_cores() -> bool:
    if _MODE == 'auto':
        return torch.is_autocast_enabled()
    return _MODE == 'bf16'
