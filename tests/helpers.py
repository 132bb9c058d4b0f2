"""Shared test helpers: golden-fixture loading and state-dict plumbing."""
import os
from collections import OrderedDict

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + '.npz'))
    return {k: z[k] for k in z.files}


def params_from_golden(g, tag):
    p = OrderedDict()
    for k, v in g.items():
        if k.startswith(tag + '.'):
            p[k[len(tag) + 1:]] = torch.from_numpy(np.asarray(v))
    return p


def t(x):
    return torch.from_numpy(np.ascontiguousarray(x))


def batch_from_golden(g):
    return {'ray_o': t(g['ray_o']), 'ray_d': t(g['ray_d']), 'depth_range': t(g['depth_range']),
            'camera': t(g['camera']), 'src_rgbs': t(g['src_rgbs']), 'src_cameras': t(g['src_cameras']),
            'rgb': t(g['gt_rgb'])}


def relerr(a, b):
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def maxabs(a, b):
    return (torch.as_tensor(a).double() - torch.as_tensor(b).double()).abs().max().item()
