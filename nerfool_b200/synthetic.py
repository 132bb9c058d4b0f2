"""Seeded synthetic scenes of the shapes BASELINE.json names.

Host-side numpy/torch only.  Produces exactly the tensors the reference's data loaders + ray sampler hand
to the hot path: ``camera`` vectors of 34 floats (``[H, W, K(4x4), c2w(4x4)]``,
/root/reference/ibrnet/data_loaders/llff_test.py:116-117), ``src_rgbs [1,V,H,W,3]``, rays built like
/root/reference/ibrnet/sample_ray.py:98-116 (integer pixel centres, ``ray_d = c2w[:3,:3] K^-1 [u,v,1]``,
un-normalised) and ``depth_range [1,2]``.
"""
from __future__ import annotations

import math

import numpy as np
import torch


def _rot(axis, deg):
    axis = np.asarray(axis, np.float64)
    axis = axis / (np.linalg.norm(axis) + 1e-12)
    a = math.radians(deg)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + math.sin(a) * K + (1 - math.cos(a)) * (K @ K)


def _camera_vector(H, W, f, c2w):
    K = np.array([[f, 0, W / 2.0, 0], [0, f, H / 2.0, 0], [0, 0, 1, 0], [0, 0, 0, 1]], np.float64)
    return np.concatenate([[H, W], K.reshape(-1), c2w.reshape(-1)]).astype(np.float32)


def _look_at(eye, target, up=(0, -1, 0)):
    """OpenCV convention (x right, y down, z forward), camera-to-world."""
    z = target - eye
    z = z / np.linalg.norm(z)
    x = np.cross(-np.asarray(up, np.float64), z)
    x = x / np.linalg.norm(x)
    y = np.cross(z, x)
    c2w = np.eye(4)
    c2w[:3, 0], c2w[:3, 1], c2w[:3, 2], c2w[:3, 3] = x, y, z, eye
    return c2w


def feature_map_size(H, W):
    """Size of the ResUNet feature maps for an HxW image: four ceil-halvings (conv1 s2, max-pool-free
    stem + three stride-2 stages) followed by two x2 upsamples, i.e. 4*ceil(n/16)-ish.  Pinned against
    the reference encoder by tests/golden/feature_sizes.npz (/root/reference/ibrnet/feature_network.py:
    231-267): 378x504 -> 96x128, 200x200 -> 52x52."""
    def halve(n, times):
        for _ in range(times):
            n = (n + 1) // 2
        return n
    return halve(H, 4) * 4, halve(W, 4) * 4


def make_scene(H=378, W=504, V=4, seed=0, kind='llff', feat_ch=32, n_targets=1):
    """Returns a dict of CPU float32 tensors:
    camera [n_targets,34], src_cameras [1,V,34], src_rgbs [1,V,H,W,3], depth_range [1,2],
    rgb [n_targets,H*W,3] (target colours), featmaps (coarse, fine) each [V,feat_ch,h',w'] (stand-ins for
    the cuDNN encoder output, smooth random), H, W."""
    rs = np.random.RandomState(seed)
    f = 0.8 * W
    if kind == 'llff':
        depth_range = np.array([[2.0, 12.0]], np.float32)
        tgt_eyes = [np.array([0.05 * rs.randn(), 0.05 * rs.randn(), 0.0]) for _ in range(n_targets)]
        tgt_c2w = []
        for eye in tgt_eyes:
            c = np.eye(4)
            c[:3, :3] = _rot(rs.randn(3), 3.0 * rs.rand() + 0.5)
            c[:3, 3] = eye
            tgt_c2w.append(c)
        src_c2w = []
        for v in range(V):
            ang = 2 * math.pi * (v + 0.37) / V
            c = np.eye(4)
            c[:3, :3] = _rot(rs.randn(3), 5.0 * rs.rand() + 0.3)
            c[:3, 3] = [0.3 * math.cos(ang), 0.3 * math.sin(ang), 0.02 * rs.randn()]
            src_c2w.append(c)
    elif kind == 'synthetic':
        depth_range = np.array([[2.0, 6.0]], np.float32)

        def on_sphere():
            d = rs.randn(3)
            d[1] = -abs(d[1]) * 0.5
            d = d / np.linalg.norm(d)
            return 4.0 * d
        base = on_sphere()
        tgt_c2w = []
        for _ in range(n_targets):
            eye = base + 0.3 * rs.randn(3)
            eye = 4.0 * eye / np.linalg.norm(eye)
            c = _look_at(eye, np.zeros(3))
            c[:3, :3] = c[:3, :3] @ _rot(rs.randn(3), 2.0 * rs.rand() + 0.3)
            tgt_c2w.append(c)
        src_c2w = []
        for v in range(V):
            eye = base + 0.9 * rs.randn(3)
            eye = 4.0 * eye / np.linalg.norm(eye)
            c = _look_at(eye, np.zeros(3))
            c[:3, :3] = c[:3, :3] @ _rot(rs.randn(3), 5.0 * rs.rand() + 0.3)
            src_c2w.append(c)
    else:
        raise ValueError(kind)

    camera = np.stack([_camera_vector(H, W, f, c) for c in tgt_c2w])
    src_cameras = np.stack([_camera_vector(H, W, f, c) for c in src_c2w])[None]

    g = torch.Generator().manual_seed(seed)
    low = torch.rand(V, 3, (H + 7) // 8 + 1, (W + 7) // 8 + 1, generator=g)
    src = torch.nn.functional.interpolate(low, size=(H, W), mode='bilinear', align_corners=True)
    src_rgbs = src.permute(0, 2, 3, 1).contiguous()[None]                      # [1,V,H,W,3]
    fh, fw = feature_map_size(H, W)
    fl = torch.randn(2, V, feat_ch, fh // 4 + 1, fw // 4 + 1, generator=g)
    feat = torch.nn.functional.interpolate(fl.reshape(2 * V, feat_ch, fh // 4 + 1, fw // 4 + 1),
                                           size=(fh, fw), mode='bilinear', align_corners=True)
    feat = feat.reshape(2, V, feat_ch, fh, fw)
    rgb = torch.rand(n_targets, H * W, 3, generator=g)
    return {
        'H': H, 'W': W,
        'camera': torch.from_numpy(camera),
        'src_cameras': torch.from_numpy(src_cameras),
        'src_rgbs': src_rgbs,
        'depth_range': torch.from_numpy(depth_range),
        'rgb': rgb,
        'featmaps': (feat[0].contiguous(), feat[1].contiguous()),
    }


def rays_for_view(camera: torch.Tensor, H: int, W: int, stride: int = 1):
    """All rays of one target view (sample_ray.py:98-116).  camera [34] or [1,34] -> ray_o, ray_d [HW,3]."""
    cam = camera.reshape(-1, 34)[:1]
    K = cam[:, 2:18].reshape(-1, 4, 4)
    c2w = cam[:, 18:34].reshape(-1, 4, 4)
    u, v = np.meshgrid(np.arange(W)[::stride], np.arange(H)[::stride])
    u = u.reshape(-1).astype(np.float32)
    v = v.reshape(-1).astype(np.float32)
    pix = torch.from_numpy(np.stack((u, v, np.ones_like(u)), axis=0))[None]
    d = (c2w[:, :3, :3].bmm(torch.inverse(K[:, :3, :3])).bmm(pix)).transpose(1, 2).reshape(-1, 3)
    o = c2w[:, :3, 3].unsqueeze(1).repeat(1, d.shape[0], 1).reshape(-1, 3)
    return o.contiguous(), d.contiguous()


def ray_batch_for(scene: dict, ray_ids, target: int = 0):
    """The dict ``render_rays`` consumes (render_ray.py:185-211 keys), for the chosen rays of one target."""
    o, d = rays_for_view(scene['camera'][target], scene['H'], scene['W'])
    ids = torch.as_tensor(ray_ids, dtype=torch.long)
    return {
        'ray_o': o[ids].contiguous(), 'ray_d': d[ids].contiguous(),
        'depth_range': scene['depth_range'],
        'camera': scene['camera'][target:target + 1],
        'rgb': scene['rgb'][target][ids].contiguous(),
        'src_rgbs': scene['src_rgbs'], 'src_cameras': scene['src_cameras'],
    }
