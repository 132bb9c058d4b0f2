"""Stub of tensorboardX.SummaryWriter: records scalars in memory (train.py logs through it)."""


class SummaryWriter:
    def __init__(self, *a, **k):
        self.scalars = []

    def add_scalar(self, tag, value, step=None, *a, **k):
        self.scalars.append((tag, float(value), step))

    def add_image(self, *a, **k):
        pass

    def close(self):
        pass
