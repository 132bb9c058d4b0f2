"""Stub of imageio (absent from this image): the reference's data loaders import it at module level."""


def imread(path, *a, **k):
    raise RuntimeError('imageio stub: no image files in the synthetic test harness (%s)' % (path,))


def imwrite(path, img, *a, **k):
    return None
