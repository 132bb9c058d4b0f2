"""Drop-in for ``ibrnet.render_ray`` (/root/reference/ibrnet/render_ray.py): ``sample_pdf``,
``sample_along_camera_ray``, ``raw2outputs``, ``render_rays`` (+ ``render_rays_hybrid``) with the
reference's signatures and return dictionaries, running on the CUDA library.

``render_rays`` takes the FUSED path (projection + gather + IBRNet + compositing per level without
materialising the [R,S,V,35] tensor) when ``projector`` / ``model.net_*`` are the nerfool_b200 types;
otherwise it composes the stand-alone operators exactly like the reference does."""
from __future__ import annotations

import os
from collections import OrderedDict

import torch

from . import ops
from .mlp_network import IBRNet
from .projection import Projector


def _unwrap(net):
    """DistributedDataParallel / DataParallel wrap the nets (model.py:78-110)."""
    return net.module if hasattr(net, 'module') and isinstance(getattr(net, 'module'), torch.nn.Module) else net


# ----------------------------------------------------------------------------------------------------
# helpers with the reference's signatures
# ----------------------------------------------------------------------------------------------------
def sample_pdf(bins, weights, N_samples, det=False):
    """
    :param bins: tensor of shape [N_rays, M+1], M is the number of bins
    :param weights: tensor of shape [N_rays, M]
    :param N_samples: number of samples along each ray
    :param det: if True, will perform deterministic sampling
    :return: [N_rays, N_samples]
    """
    w_in = weights.detach().clone()   # the kernel applies the 1e-5 floor itself (same fp32 rounding)
    weights += 1e-5                   # the reference mutates its argument (render_ray.py:34): keep that effect
    u = _uniforms(bins.shape[0], N_samples, det, bins.device)
    return ops.sample_pdf_op(bins, w_in, u)


_linspace_cache = {}


def _uniforms(R, n, det, device):
    if det:
        # torch.linspace on CPU then copy: the values the reference's CPU path sees (render_ray.py:42); the device
        # copy is cached per (n, device) -- no host->device copy per step (and none inside a CUDA-graph capture)
        key = (int(n), str(device))
        u = _linspace_cache.get(key)
        if u is None:
            u = _linspace_cache[key] = torch.linspace(0., 1., n).to(device)
        return u
    return torch.rand(R, n, device=device)


def sample_along_camera_ray(ray_o, ray_d, depth_range, N_samples, inv_uniform=False, det=False):
    """
    :return: pts [N_rays, N_samples, 3], z_vals [N_rays, N_samples]   (render_ray.py:73-116)
    """
    near, far = ops.depth_range_pair(depth_range)
    assert near > 0 and far > 0 and far > near
    R = ray_d.shape[0]
    t_rand = None if det else torch.rand(R, N_samples, device=ray_d.device)
    z_vals = ops.coarse_depths(R, N_samples, near, far, inv_uniform, t_rand, ray_d.device)
    pts = z_vals.unsqueeze(2) * ray_d.unsqueeze(1) + ray_o.unsqueeze(1)
    return pts, z_vals


def raw2outputs(raw, z_vals, mask, white_bkgd=False, geo_noise=None):
    """
    :param raw: raw network output; tensor of shape [N_rays, N_samples, 4]
    :param z_vals: depth of point samples along rays; tensor of shape [N_rays, N_samples]
    :param mask: [N_rays, N_samples] (pixel mask: at least two observations)
    :return: OrderedDict rgb, depth, weights, mask, alpha, z_vals   (render_ray.py:123-170)
    """
    if geo_noise is not None and geo_noise > 0:
        noise = torch.zeros_like(raw)
        noise[..., 3] = torch.randn_like(raw[..., 3]) * geo_noise
        raw = raw + noise
    rgb, depth, weights, alpha, ray_mask = ops.Composite.apply(raw, z_vals, mask, bool(white_bkgd))
    return OrderedDict([('rgb', rgb), ('depth', depth), ('weights', weights), ('mask', ray_mask),
                        ('alpha', alpha), ('z_vals', z_vals)])


def _fine_z(z_vals, weights, N_importance, inv_uniform, det):
    """render_ray.py:216-238 in one kernel (mid-points, flip, sample_pdf, inversion, merge-sort)."""
    u = _uniforms(z_vals.shape[0], N_importance, det, z_vals.device)
    return ops.fine_depths(z_vals, weights, u, inv_uniform)


# ----------------------------------------------------------------------------------------------------
# render_rays
# ----------------------------------------------------------------------------------------------------
def _fusable(model, projector):
    if os.environ.get('NFB_FUSED', '1') == '0':
        return False
    nc = _unwrap(model.net_coarse)
    nf = _unwrap(model.net_fine) if getattr(model, 'net_fine', None) is not None else None
    if not (isinstance(projector, Projector) and isinstance(nc, IBRNet) and (nf is None or isinstance(nf, IBRNet))):
        return False
    # A DistributedDataParallel / DataParallel wrapper (model.py:78-110) must see its own forward() when the parameters
    # train: its gradient-reduction hooks are armed there.  The fused path calls the unwrapped module, so in that case
    # the composed path (Projector.compute -> wrapper(rgb_feat, ray_diff, mask) -> raw2outputs) is used instead.
    for net, inner in ((model.net_coarse, nc), (getattr(model, 'net_fine', None), nf)):
        if inner is not None and net is not inner and inner.training and torch.is_grad_enabled() and \
                any(p.requires_grad for p in inner.parameters()):
            return False
    return True


def render_rays(ray_batch, model, featmaps, projector, N_samples, inv_uniform=False, N_importance=0, det=False,
                white_bkgd=False, args=None, src_ray_batch=None, geo_noise=None):
    """
    :param ray_batch: {'ray_o': [N_rays, 3] , 'ray_d': [N_rays, 3], 'depth_range', 'camera', 'src_rgbs',
                       'src_cameras'}
    :param model:  object with .net_coarse / .net_fine
    :return: {'outputs_coarse': {}, 'outputs_fine': {}}          (render_ray.py:173-256)
    """
    ret = {'outputs_coarse': None, 'outputs_fine': None}
    src = ray_batch if src_ray_batch is None else src_ray_batch
    ray_o, ray_d = ray_batch['ray_o'], ray_batch['ray_d']
    near, far = ops.depth_range_pair(ray_batch['depth_range'])
    assert near > 0 and far > 0 and far > near
    R = ray_d.shape[0]
    dev = ray_d.device
    if 'nfb_coarse_z' in ray_batch and det:
        z_vals = ray_batch['nfb_coarse_z']        # caller-owned static buffer (attack.GraphedPGDStep)
    else:
        t_rand = None if det else torch.rand(R, N_samples, device=dev)
        z_vals = ops.coarse_depths(R, N_samples, near, far, inv_uniform, t_rand, dev)

    fused = _fusable(model, projector)
    noise = float(geo_noise) if (geo_noise is not None and geo_noise > 0) else 0.0
    if fused:
        src_rgbs, src_cams = src['src_rgbs'], src['src_cameras']
        assert src_rgbs.shape[0] == 1 and src_cams.shape[0] == 1 and ray_batch['camera'].shape[0] == 1, \
            'only support batch_size=1 for now'
        H, W = int(src_rgbs.shape[2]), int(src_rgbs.shape[3])
        cam = ray_batch.get('nfb_camera_block')     # caller-owned static block (attack.GraphedPGDStep)
        if cam is None:
            cam = ops.camera_block(src_cams[0], ray_batch['camera'][0], dev, H, W)

        def level(net, fmap, z):
            net = _unwrap(net)
            rgb, depth, weights, alpha, mask = ops.RenderLevel.apply(
                fmap, src_rgbs[0], ray_o, ray_d, z, cam, net.param_blob(), net.pos_encoding[0], H, W,
                bool(net.anti_alias_pooling), bool(white_bkgd), noise)
            return OrderedDict([('rgb', rgb), ('depth', depth), ('weights', weights), ('mask', mask),
                                ('alpha', alpha), ('z_vals', z)])
    else:
        def level(net, fmap, z):
            pts = z.unsqueeze(2) * ray_d.unsqueeze(1) + ray_o.unsqueeze(1)
            rgb_feat, ray_diff, mask = projector.compute(pts, ray_batch['camera'], src['src_rgbs'],
                                                         src['src_cameras'], featmaps=fmap)
            pixel_mask = mask[..., 0].sum(dim=2) > 1
            raw = net(rgb_feat, ray_diff, mask)
            return raw2outputs(raw, z, pixel_mask, white_bkgd=white_bkgd, geo_noise=geo_noise)

    ret['outputs_coarse'] = level(model.net_coarse, featmaps[0], z_vals)
    if N_importance > 0:
        assert model.net_fine is not None
        weights = ret['outputs_coarse']['weights'].detach()
        z_fine = _fine_z(z_vals, weights, N_importance, inv_uniform, det)
        ret['outputs_fine'] = level(model.net_fine, featmaps[1], z_fine)
    return ret


def render_rays_hybrid(ray_batch, model, featmaps, projector, N_samples, inv_uniform=False, N_importance=0, det=False,
                       white_bkgd=False, args=None, src_ray_batch=None, featmaps_clean=None):
    """
    Clean / adversarial mixing ablation (render_ray.py:261-390): both feature-map sets are projected and aggregated at
    the same points; colour and density are each taken from the clean or the adversarial pass (args.use_clean_color,
    args.use_clean_density) before compositing.
    :return: {'outputs_coarse': {}, 'outputs_fine': {}}
    """
    ret = {'outputs_coarse': None, 'outputs_fine': None}
    src = ray_batch if src_ray_batch is None else src_ray_batch
    ray_o, ray_d = ray_batch['ray_o'], ray_batch['ray_d']
    _, z_vals = sample_along_camera_ray(ray_o=ray_o, ray_d=ray_d, depth_range=ray_batch['depth_range'],
                                        N_samples=N_samples, inv_uniform=inv_uniform, det=det)

    def level(net, fmap_adv, fmap_clean, z):
        pts = z.unsqueeze(2) * ray_d.unsqueeze(1) + ray_o.unsqueeze(1)
        raws = []
        pixel_mask = None
        for fmap in (fmap_adv, fmap_clean):
            rgb_feat, ray_diff, mask = projector.compute(pts, ray_batch['camera'], src['src_rgbs'], src['src_cameras'],
                                                         featmaps=fmap)
            if pixel_mask is None:                       # the reference composites with the adversarial pass's mask
                pixel_mask = mask[..., 0].sum(dim=2) > 1
            raws.append(net(rgb_feat, ray_diff, mask))
        raw_adv, raw_clean = raws
        color = raw_clean[:, :, :3] if args.use_clean_color else raw_adv[:, :, :3]
        sigma = raw_clean[:, :, 3:4] if args.use_clean_density else raw_adv[:, :, 3:4]
        return raw2outputs(torch.cat([color, sigma], dim=2), z, pixel_mask, white_bkgd=white_bkgd)

    ret['outputs_coarse'] = level(model.net_coarse, featmaps[0], featmaps_clean[0], z_vals)
    if N_importance > 0:
        assert model.net_fine is not None
        weights = ret['outputs_coarse']['weights'].clone().detach()
        z_fine = _fine_z(z_vals, weights, N_importance, inv_uniform, det)
        ret['outputs_fine'] = level(model.net_fine, featmaps[1], featmaps_clean[1], z_fine)
    return ret
