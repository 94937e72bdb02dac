"""Stub (test infrastructure): matplotlib is imported by the reference's attention-map plotting helper only (src/utils.py:11-14)."""
