"""Stub: the reference imports matplotlib at module level (train.py:9, utils.py:8) but never plots on the hot path."""
