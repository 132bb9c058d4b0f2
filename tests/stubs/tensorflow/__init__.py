"""Stub of tensorflow (eval_adv.py builds an LPIPS-TF session in its __main__ block only)."""


def __getattr__(name):
    raise RuntimeError('tensorflow stub: LPIPS-TF is out of scope (%s)' % name)
