"""Stub of matplotlib (utils.py imports it for colour-bar plots, which the test harness never draws)."""
from . import cm, figure  # noqa: F401


def use(*a, **k):
    pass


class _Unavailable:
    def __getattr__(self, name):
        raise RuntimeError('matplotlib stub: plotting is not available in the test harness')


colors = _Unavailable()
colorbar = _Unavailable()
ticker = _Unavailable()
