"""Drop-in for ``ibrnet.projection.Projector`` (/root/reference/ibrnet/projection.py:20-132).

Same constructor and ``compute`` signature / return values; the projection, in-frustum mask, ray_diff
and both bilinear gathers run in one CUDA kernel (``nfb_project_gather_fwd``), the backward is the
scatter kernel ``nfb_project_gather_bwd``."""
from __future__ import annotations

import torch

from . import ops


class Projector:
    def __init__(self, device):
        self.device = device

    def compute(self, xyz, query_camera, train_imgs, train_cameras, featmaps):
        """
        :param xyz: [n_rays, n_samples, 3]
        :param query_camera: [1, 34], 34 = img_size(2) + intrinsics(16) + extrinsics(16)
        :param train_imgs: [1, n_views, h, w, 3]
        :param train_cameras: [1, n_views, 34]
        :param featmaps: [n_views, d, h', w']
        :return: rgb_feat_sampled [n_rays, n_samples, n_views, 3+d], ray_diff [.., n_views, 4],
                 mask [.., n_views, 1]
        """
        assert (train_imgs.shape[0] == 1) and (train_cameras.shape[0] == 1) and (query_camera.shape[0] == 1), \
            'only support batch_size=1 for now'
        H, W = int(train_imgs.shape[2]), int(train_imgs.shape[3])
        cams = train_cameras[0]
        # the reference normalises with h, w read from the camera vector (projection.py:112); camera_block checks (once per
        # cached camera tensor) that they equal the image size the kernels use
        cam = ops.camera_block(cams, query_camera[0], xyz.device, H, W)
        return ops.ProjectGather.apply(xyz, train_imgs[0], featmaps, cam, H, W)
