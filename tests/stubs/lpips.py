"""Stub of lpips (eval/gnt drivers): eval/gnt/utils.py builds two LPIPS nets at import time; using one raises."""


class LPIPS:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        raise RuntimeError('lpips stub: LPIPS is out of scope')

    forward = __call__
