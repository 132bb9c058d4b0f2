"""Generate tests/golden/gnt_d2.npz by running the UNMODIFIED reference GNT (imported from /root/reference) on seeded
inputs.  Build container only; the fixture is committed.  Usage:  python oracle/make_golden_gnt.py

Reference code exercised: gnt/transformer_network.py:205-309 (GNT, trans_depth=2, netwidth=64, ret_alpha=True, eval
mode) on Projector.compute outputs of a synthetic scene (gnt/projection.py == ibrnet/projection.py up to formatting).
"""
import os
import sys
import types

import numpy as np
import torch

sys.dont_write_bytecode = True
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(1, '/root/reference')

from gnt.transformer_network import GNT            # noqa: E402  (reference)
from gnt.projection import Projector               # noqa: E402  (reference)
from gnt.render_ray import sample_along_camera_ray  # noqa: E402  (reference)

from nerfool_b200.synthetic import make_scene, ray_batch_for   # noqa: E402


def main():
    depth, V, R, S, H, W, seed = 2, 5, 12, 24, 96, 128, 7
    scene = make_scene(H, W, V, seed=seed, kind='llff')
    ids = np.sort(np.random.RandomState(seed + 1).choice(H * W, R, replace=False))
    batch = ray_batch_for(scene, ids)
    torch.manual_seed(seed)
    net = GNT(types.SimpleNamespace(netwidth=64, trans_depth=depth), in_feat_ch=32, posenc_dim=63, viewenc_dim=63, ret_alpha=True)
    with torch.no_grad():
        for name, prm in net.named_parameters():
            if name.endswith('.bias') or 'norm' in name:
                prm += 0.05 * torch.randn_like(prm)
    net.eval()
    pts, z = sample_along_camera_ray(batch['ray_o'], batch['ray_d'], batch['depth_range'], S, inv_uniform=True, det=True)
    rgb_feat, ray_diff, mask = Projector(device='cpu').compute(pts, batch['camera'], batch['src_rgbs'], batch['src_cameras'],
                                                               featmaps=scene['featmaps'][0])
    with torch.no_grad():
        out = net(rgb_feat, ray_diff, mask, pts, batch['ray_d'])
    d = {'p.' + k: v.detach().numpy().copy() for k, v in net.state_dict().items()}
    d.update(depth=np.int64(depth), rgb_feat=rgb_feat.numpy(), ray_diff=ray_diff.numpy(), mask=mask.numpy(), pts=pts.numpy(),
             ray_d=batch['ray_d'].numpy(), z_vals=z.numpy(), out=out.numpy())
    path = os.path.join(REPO, 'tests', 'golden', 'gnt_d2.npz')
    np.savez_compressed(path, **d)
    print('wrote', path, os.path.getsize(path) // 1024, 'KiB; out', out.shape)


if __name__ == '__main__':
    main()
