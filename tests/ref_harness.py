"""Harness that runs the reference's own DRIVERS, unmodified, on top of ``dropin/``.

TEST INFRASTRUCTURE.  The reference checkout is imported from ``/root/reference`` when present (build container) or from
its byte-for-byte staged copy ``baseline/_ref`` (``oracle/stage_reference.py``; git-ignored, travels to the GPU box).
``sys.path`` order: ``tests/stubs`` (imageio / matplotlib / configargparse / tensorboardX / tensorflow / lpips stand-ins)
-> ``<repo>/dropin`` (ibrnet.projection / mlp_network / render_ray / render_image -> nerfool_b200) -> ``<repo>`` ->
reference root (config.py, utils.py, train.py, ibrnet/*, gnt/*) -> reference ``eval/ibrnet`` (eval_adv.py, geo_interp.py,
pc_grad.py).  A synthetic dataset class is registered in the reference's own ``dataset_dict``.
"""
from __future__ import annotations

import importlib
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from oracle.stage_reference import staged_root  # noqa: E402

_GENERIC = ('ibrnet', 'gnt', 'utils', 'config', 'train', 'eval_adv', 'geo_interp', 'pc_grad', 'imageio', 'matplotlib',
            'configargparse', 'tensorboardX', 'tensorflow', 'lpips_tensorflow', 'lpips')


def reference_available() -> bool:
    return staged_root() is not None


def setup_paths(kind: str = 'ibrnet') -> str:
    """Arrange sys.path / sys.modules so that ``import train`` / ``import eval_adv`` load the reference's files with the
    hot-path modules overlaid by dropin/.  Returns the reference root."""
    root = staged_root()
    if root is None:
        raise RuntimeError('no reference checkout (/root/reference) and no staged copy (baseline/_ref): run oracle/stage_reference.py')
    os.environ['NERFOOL_REFERENCE_ROOT'] = root
    for name in list(sys.modules):
        if name.split('.')[0] in _GENERIC:
            del sys.modules[name]
    # the driver's own directory first, as when it is run from there (eval/gnt/ holds the GNT path's own config.py / train.py / utils.py)
    want = [os.path.join(REPO, 'tests', 'stubs'), os.path.join(REPO, 'dropin'), REPO, os.path.join(root, 'eval', kind), root]
    sys.path[:] = want + [p for p in sys.path if p not in want]
    importlib.invalidate_caches()
    return root


class SyntheticSceneDataset(torch.utils.data.Dataset):
    """Stands where ``LLFFTestDataset`` stands (ibrnet/data_loaders/llff_test.py:102-204): one item per target view with the same
    keys, dtypes and shapes (``rgb [H,W,3]``, ``camera [34]``, ``rgb_path``, ``src_rgbs [V,H,W,3]``, ``src_cameras [V,34]``,
    ``depth_range [2]``), from the seeded scene generator of nerfool_b200.synthetic."""
    H, W, V, N_TARGETS, SEED = 96, 128, 4, 3, 0

    def __init__(self, args, mode, scenes=(), **kwargs):
        from nerfool_b200.synthetic import make_scene
        self.mode = mode
        self.scene = make_scene(self.H, self.W, self.V, seed=self.SEED, kind='llff', n_targets=self.N_TARGETS)

    def __len__(self):
        return self.N_TARGETS

    def __getitem__(self, idx):
        s = self.scene
        idx = idx % self.N_TARGETS
        return {'rgb': s['rgb'][idx].reshape(self.H, self.W, 3).clone(), 'camera': s['camera'][idx].clone(),
                'rgb_path': f'synthetic_{idx:03d}.png', 'src_rgbs': s['src_rgbs'][0].clone(),
                'src_cameras': s['src_cameras'][0].clone(), 'depth_range': s['depth_range'][0].clone()}


def register_dataset(name: str = 'synthetic_b200', **shape):
    """Add the synthetic dataset to the reference's own ``dataset_dict`` (ibrnet/data_loaders/__init__.py:28-37)."""
    kind = shape.pop('kind', 'ibrnet')
    dataset_dict = importlib.import_module(kind + '.data_loaders').dataset_dict
    cls = type('SyntheticSceneDataset_' + name, (SyntheticSceneDataset,), dict(shape))
    dataset_dict[name] = cls
    return cls


def parse_args(root: str, extra: list[str], config: str = 'configs/ibrnet/eval_llff.txt'):
    """The reference's own ``config.config_parser()`` on one of its own config files (+ overrides)."""
    import config as ref_config
    parser = ref_config.config_parser()
    argv = ['--config', os.path.join(root, config)] + list(extra)
    return parser.parse_args(argv)


def seed_everything(seed: int = 0, kind: str = 'ibrnet'):
    torch.manual_seed(seed)
    np.random.seed(seed)
    sample_ray = importlib.import_module(kind + '.sample_ray')
    sample_ray.rng.seed(234)          # the module-level RandomState(234) of sample_ray.py:20
