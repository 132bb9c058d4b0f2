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


# ---- BASELINE-shaped fixtures (oracle/make_golden_baseline.py): inputs are regenerated from the seed, digests checked ----
def _sha(x):
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(x.detach().cpu().numpy()).tobytes()).hexdigest()


def scene_from_golden(g):
    """Rebuild the synthetic scene a base_* fixture was made from; returns (scene, batch, inputs_bit_identical)."""
    import sys
    sys.path.insert(0, os.path.dirname(GOLDEN.rstrip('/')).rsplit('/tests', 1)[0])
    from nerfool_b200.synthetic import make_scene, ray_batch_for
    scene = make_scene(int(g['H']), int(g['W']), int(g['V']), seed=int(g['seed']), kind=str(g['kind']))
    same = all(_sha(x) == str(g[k]) for k, x in (('sha_src_rgbs', scene['src_rgbs']), ('sha_feat_c', scene['featmaps'][0]),
                                                  ('sha_feat_f', scene['featmaps'][1]), ('sha_camera', scene['camera']),
                                                  ('sha_src_cameras', scene['src_cameras']), ('sha_rgb', scene['rgb'])))
    return scene, ray_batch_for(scene, g['ray_ids']), same


def sampled_grad_relerr(grad_nchw, g, tag):
    """Relative error of a feature-map gradient [V,32,h,w] against the fixture's sampled texels (rows of the channel-last
    gradient), normalised by the norm of the reference rows."""
    rows = torch.as_tensor(grad_nchw).detach().cpu().permute(0, 2, 3, 1).reshape(-1, 32)[torch.from_numpy(g[f'd_feat_{tag}_idx']).long()]
    ref = torch.from_numpy(g[f'd_feat_{tag}_val'])
    return relerr(rows, ref)


def report(msg):
    """Measured parity numbers go to the pytest log (run with -rP / -s to see them; also appended to
    gpurun_out/parity_report.txt when that directory exists)."""
    print('[parity] ' + msg)
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out')
    if os.path.isdir(d):
        with open(os.path.join(d, 'parity_report.txt'), 'a') as f:
            f.write(msg + '\n')
