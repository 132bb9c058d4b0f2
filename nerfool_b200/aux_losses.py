"""Auxiliary attack losses next to the hot path, on the device.

``calc_depth_smooth_loss`` (/root/reference/eval/ibrnet/eval_adv.py:32-48) and ``calc_depth_var``
(/root/reference/train.py:329-340) are plain tensor expressions of ``render_rays`` outputs: they run unchanged on the CUDA
tensors the fused path returns, and their gradients reach the feature maps through ``d depth`` / ``d weights`` of the
compositing backward (``nfb_composite_bwd``).  They are restated here so that callers outside the reference have them.

``forward_warp`` (eval_adv.py:97-197) is different: the reference walks all H x W pixels in a Python loop on CPU tensors -- once
per PGD iteration when the depth- or camera-consistency loss is on -- which costs seconds and dominates a fast renderer.  Here the
z-buffer splat is ``nfb_forward_warp`` (two streaming passes with 64-bit atomicMin keys); same signature, same return values, on
the device.  The reference defines its loop inside ``eval_adv.py``; the one-line hook for the unmodified script is
``eval_adv.forward_warp = nerfool_b200.aux_losses.forward_warp``."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._lib import call, ptr, stream_ptr


def calc_depth_smooth_loss(ret, patch_size, loss_type='l2'):
    """eval_adv.py:32-48: squared (or absolute) forward differences of the depth inside each patch_size x patch_size ray patch."""
    depth = ret['depth'].reshape([-1, patch_size, patch_size])
    v00, v01, v10 = depth[:, :-1, :-1], depth[:, :-1, 1:], depth[:, 1:, :-1]
    if loss_type == 'l2':
        loss = ((v00 - v01) ** 2) + ((v00 - v10) ** 2)
    elif loss_type == 'l1':
        loss = torch.abs(v00 - v01) + torch.abs(v00 - v10)
    else:
        raise ValueError('Not supported loss type.')
    return loss.sum()


def calc_depth_var(ret):
    """train.py:329-340: mean over rays of the weight-normalised depth variance (NaN rays -- zero total weight -- dropped)."""
    depth, weights, z_vals = ret['depth'], ret['weights'], ret['z_vals']
    var = torch.sum(weights * (z_vals - depth.unsqueeze(dim=1)) ** 2, dim=1) / torch.sum(weights, dim=1)
    var = torch.masked_select(var, ~torch.isnan(var))
    return torch.mean(var)


def project_with_depth(depth_ref, intrinsics_ref, extrinsics_ref, intrinsics_src, extrinsics_src):
    """eval_adv.py:62-94, the reference's own tensor expressions (they already run on the device there)."""
    width, height = depth_ref.shape[2], depth_ref.shape[1]
    batchsize = depth_ref.shape[0]
    y_ref, x_ref = torch.meshgrid([torch.arange(0, height, dtype=torch.float32, device=depth_ref.device),
                                   torch.arange(0, width, dtype=torch.float32, device=depth_ref.device)], indexing='ij')
    y_ref, x_ref = y_ref.contiguous().view(height * width), x_ref.contiguous().view(height * width)
    pts = torch.stack((x_ref, y_ref, torch.ones_like(x_ref))).unsqueeze(0) * (depth_ref.view(batchsize, -1).unsqueeze(1))
    xyz_ref = torch.matmul(torch.inverse(intrinsics_ref), pts)
    xyz_src = torch.matmul(torch.matmul(torch.inverse(extrinsics_src), extrinsics_ref),
                           torch.cat((xyz_ref, torch.ones_like(x_ref.unsqueeze(0)).repeat(batchsize, 1, 1)), dim=1))[:, :3, :]
    K_xyz_src = torch.matmul(intrinsics_src, xyz_src)
    depth_src = K_xyz_src[:, 2:3, :]
    xy_src = K_xyz_src[:, :2, :] / (K_xyz_src[:, 2:3, :] + 1e-9)
    x_src = xy_src[:, 0, :].view([batchsize, height, width])
    y_src = xy_src[:, 1, :].view([batchsize, height, width])
    return x_src, y_src, depth_src


def _splat(H, W, x_res, y_res, depth_src, rgb_ref, allowed, sources):
    dev = depth_src.device
    n = H * W
    new = torch.empty(n, 3, device=dev, dtype=torch.float32)
    new_depth = torch.empty(n, device=dev, dtype=torch.float32)
    keys = torch.empty(n, device=dev, dtype=torch.int64)
    flag = torch.empty(1, device=dev, dtype=torch.int32)
    n_src = 0 if sources is None else int(sources.numel())
    args = (H, W, ptr(x_res), ptr(y_res), ptr(depth_src), ptr(rgb_ref), ptr(allowed), ptr(sources), n_src, ptr(new), ptr(new_depth),
            ptr(keys), ptr(flag))
    with torch.cuda.device(dev):
        call('nfb_forward_warp', *args, 0, stream_ptr(dev))
        if int(flag.item()):        # a depth <= 0 / NaN: 0 doubles as the reference's "empty" marker -> its exact sequential loop
            call('nfb_forward_warp', *args, 1, stream_ptr(dev))
    return new.view(H, W, 3), new_depth.view(H, W)


def forward_warp(selected_inds, rgb_ref, depth_ref, intrinsics_ref, extrinsics_ref, intrinsics_src, extrinsics_src, src2tar=True,
                 derive_full_image=False, cpu_speedup=True):
    """
    eval_adv.py:97-197.  selected_inds: [Num_Sampled_Rays] pixel indices (numpy / list / tensor); rgb_ref [H, W, 3];
    depth_ref [1, H, W]; intrinsics [3, 3]; extrinsics [4, 4].  Returns (new [H,W,3], new_depth [H,W], rgb_proj [N,3],
    depth_proj [N]) and, for src2tar=False, additionally the list selected_inds_new -- all tensors on depth_ref's device
    (``cpu_speedup`` is accepted and ignored: nothing runs on the CPU).
    """
    _lib.require_cuda(depth_ref, rgb_ref)
    if depth_ref.dim() == 2:
        depth_ref = depth_ref.unsqueeze(0)
    assert depth_ref.shape[0] == 1, 'assume batch_size=1 (eval_adv.py:119)'
    dev = depth_ref.device
    x_res, y_res, depth_src = project_with_depth(depth_ref, intrinsics_ref.to(dev), extrinsics_ref.to(dev), intrinsics_src.to(dev),
                                                 extrinsics_src.to(dev))
    width, height = depth_ref.shape[2], depth_ref.shape[1]
    if tuple(rgb_ref.shape[:2]) != (height, width):
        raise IndexError(f'forward_warp indexes rgb_ref {tuple(rgb_ref.shape)} with the pixels of a {height} x {width} depth map')
    depth_src = depth_src.reshape(-1).float().contiguous()
    y_i = torch.clamp(y_res, 0, height - 1).to(torch.long).reshape(-1).to(torch.int32).contiguous()
    x_i = torch.clamp(x_res, 0, width - 1).to(torch.long).reshape(-1).to(torch.int32).contiguous()
    rgb = rgb_ref.reshape(-1, rgb_ref.shape[-1])[:, :3].float().contiguous()
    sel = torch.as_tensor(np.asarray(selected_inds) if not torch.is_tensor(selected_inds) else selected_inds).to(dev).long().reshape(-1)
    if derive_full_image:
        new, new_depth = _splat(height, width, x_i, y_i, depth_src, rgb, None, None)
        idx = sel
    elif src2tar:
        allowed = torch.zeros(height * width, device=dev, dtype=torch.uint8)
        allowed[sel] = 1
        new, new_depth = _splat(height, width, x_i, y_i, depth_src, rgb, allowed, None)
        idx = sel
    else:
        new, new_depth = _splat(height, width, x_i, y_i, depth_src, rgb, None, sel.to(torch.int32).contiguous())
        idx = y_i.long()[sel] * width + x_i.long()[sel]
    depth_proj = new_depth.reshape(-1)[idx]
    rgb_proj = new.reshape(-1, 3)[idx]
    if not derive_full_image and not src2tar:
        return new, new_depth, rgb_proj, depth_proj, idx.tolist()
    return new, new_depth, rgb_proj, depth_proj
