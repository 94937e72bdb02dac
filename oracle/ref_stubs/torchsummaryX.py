"""Stub (test infrastructure): the reference imports torchsummaryX.summary for a model printout only (src/models.py:9, src/train.py:19)."""


def summary(*args, **kwargs):
    return None
