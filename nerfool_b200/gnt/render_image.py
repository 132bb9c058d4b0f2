"""Drop-in for ``gnt.render_image.render_single_image`` (/root/reference/gnt/render_image.py:6-135): the reference's
signature and return value, device-resident -- chunks write into frame-sized device buffers and every key is copied to
the host once per frame instead of once per chunk (the reference synchronises the device for every chunk and key)."""
from __future__ import annotations

import os
from collections import OrderedDict

import torch

from .render_ray import render_rays

_SHARED_KEYS = ('camera', 'depth_range', 'src_rgbs', 'src_cameras')


def _to_host(t: torch.Tensor) -> torch.Tensor:
    """Device -> host copy into PINNED memory, asynchronous on the current stream (the caller synchronises once per
    frame).  A frame is ~0.4 GB of outputs: pageable `.cpu()` copies run at 2-3 GB/s and took 3 x the rendering itself;
    PyTorch's caching host allocator makes the pinned buffers free after the first frame."""
    if not t.is_cuda:
        return t
    out = torch.empty(t.shape, dtype=t.dtype, device='cpu', pin_memory=True)
    out.copy_(t, non_blocking=True)
    return out


def render_single_image(ray_sampler, ray_batch, model, projector, chunk_size, N_samples, inv_uniform=False, N_importance=0,
                        det=False, white_bkgd=False, render_stride=1, featmaps=None, ret_alpha=False, single_net=False,
                        args=None, src_ray_batch=None, featmaps_clean=None):
    """
    :param ray_sampler: RaySamplingSingleImage for this view
    :param chunk_size: number of rays in a chunk
    :param ret_alpha: the network also returns the attention-derived sample weights (-> depth)
    :return: {'outputs_coarse': {'rgb': [H, W, 3], 'depth': [H, W], 'weights': [H, W, S]}, 'outputs_fine': ...}  (CPU tensors)
    """
    if args is not None and (getattr(args, 'use_clean_color', False) or getattr(args, 'use_clean_density', False)):
        raise NotImplementedError('nerfool_b200.gnt: the clean/adversarial mixing ablation (render_rays_hybrid, '
                                  'gnt/render_ray.py:281-390) is not built for the GNT path')
    N_rays = ray_batch['ray_o'].shape[0]
    # Rays are independent, so the chunk size only bounds memory.  The reference's default (4096 rays, sized for eager
    # PyTorch's saved tensors) leaves a B200 launch-bound; without autograd a chunk needs < 100 KB per ray, so render in
    # chunks of at least NFB_RENDER_CHUNK rays (default 32768; identical output).
    if not torch.is_grad_enabled():
        chunk_size = max(int(chunk_size), int(os.environ.get('NFB_RENDER_CHUNK', '32768')))
    buf = {'outputs_coarse': None, 'outputs_fine': None}
    for i in range(0, N_rays, chunk_size):
        chunk = OrderedDict()
        for k in ray_batch:
            if k in _SHARED_KEYS:
                chunk[k] = ray_batch[k]
            elif ray_batch[k] is not None:
                chunk[k] = ray_batch[k][i:i + chunk_size]
            else:
                chunk[k] = None
        ret = render_rays(chunk, model, featmaps, projector=projector, N_samples=N_samples, inv_uniform=inv_uniform,
                          N_importance=N_importance, det=det, white_bkgd=white_bkgd, ret_alpha=ret_alpha,
                          single_net=single_net, args=args, src_ray_batch=src_ray_batch)
        for lvl in ('outputs_coarse', 'outputs_fine'):
            out = ret[lvl]
            if out is None:
                continue
            if buf[lvl] is None:
                buf[lvl] = OrderedDict((k, torch.empty((N_rays,) + tuple(v.shape[1:]), dtype=v.dtype, device=v.device))
                                       for k, v in out.items() if v is not None)
            for k, v in out.items():
                if v is not None:
                    buf[lvl][k][i:i + v.shape[0]] = v.detach()
    Hs = len(range(0, ray_sampler.H, render_stride))
    Ws = len(range(0, ray_sampler.W, render_stride))
    all_ret = OrderedDict([('outputs_coarse', OrderedDict()), ('outputs_fine', OrderedDict())])
    for lvl in ('outputs_coarse', 'outputs_fine'):
        if buf[lvl] is None:
            all_ret[lvl] = None
            continue
        for k, v in buf[lvl].items():
            if k == 'random_sigma':
                continue
            all_ret[lvl][k] = _to_host(v.reshape(Hs, Ws, -1).squeeze())     # one device -> host copy per key
    torch.cuda.current_stream().synchronize()      # all asynchronous device -> host copies have landed
    return all_ret
