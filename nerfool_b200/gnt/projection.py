"""Drop-in for ``gnt.projection.Projector`` (/root/reference/gnt/projection.py:5-132).

Same arithmetic as ``ibrnet.projection.Projector`` with one difference the attack relies on: the source cameras stay in the autograd
graph (the IBRNet projector detaches them), so ``eval/gnt/eval_adv.py --perturb_camera`` (:749-869) can optimise source rotations and
translations.  When ``train_cameras`` requires a gradient the call goes through ``ops.ProjectGatherCam`` (grid gradient of both
gathers in CUDA, camera chain in torch); otherwise through the same function as the IBRNet projector."""
from __future__ import annotations

import torch

from .. import ops
from ..projection import Projector as _IBRNetProjector


class Projector(_IBRNetProjector):
    def compute(self, xyz, query_camera, train_imgs, train_cameras, featmaps):
        if not (torch.is_grad_enabled() and train_cameras.requires_grad):
            return super().compute(xyz, query_camera, train_imgs, train_cameras, featmaps)
        assert (train_imgs.shape[0] == 1) and (train_cameras.shape[0] == 1) and (query_camera.shape[0] == 1), \
            'only support batch_size=1 for now'
        H, W = int(train_imgs.shape[2]), int(train_imgs.shape[3])
        return ops.ProjectGatherCam.apply(xyz, train_imgs[0], featmaps, train_cameras[0], query_camera[0], H, W)
