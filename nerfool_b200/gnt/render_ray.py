"""Drop-in for ``gnt.render_ray.render_rays`` (/root/reference/gnt/render_ray.py:196-279): same signature and return
dict; depths, projection + gather, importance sampling and the GNT network all run in the CUDA library."""
from __future__ import annotations

import torch

from .. import ops
from ..render_ray import _fine_z, _uniforms    # noqa: F401  (shared with the IBRNet path)


def render_rays(ray_batch, model, featmaps, projector, N_samples, inv_uniform=False, N_importance=0, det=False,
                white_bkgd=False, ret_alpha=False, single_net=True, args=None, src_ray_batch=None, geo_noise=None):
    """
    :param ray_batch: {'ray_o': [N_rays, 3] , 'ray_d': [N_rays, 3], 'depth_range', 'camera', 'src_rgbs', 'src_cameras'}
    :param model: object with .net_coarse (and .net_fine unless single_net)
    :param ret_alpha: the network also returns the attention-derived sample weights (-> depth)
    :return: {'outputs_coarse': {'rgb', 'weights', 'depth'}, 'outputs_fine': ... or None}
    """
    ret = {'outputs_coarse': None, 'outputs_fine': None}
    src = ray_batch if src_ray_batch is None else src_ray_batch
    ray_o, ray_d = ray_batch['ray_o'], ray_batch['ray_d']
    near, far = ops.depth_range_pair(ray_batch['depth_range'])
    assert near > 0 and far > 0 and far > near
    R, dev = ray_d.shape[0], ray_d.device
    t_rand = None if det else torch.rand(R, N_samples, device=dev)
    z_vals = ops.coarse_depths(R, N_samples, near, far, inv_uniform, t_rand, dev)

    def level(net, fmap, z):
        pts = z.unsqueeze(2) * ray_d.unsqueeze(1) + ray_o.unsqueeze(1)
        rgb_feat, ray_diff, mask = projector.compute(pts, ray_batch['camera'], src['src_rgbs'], src['src_cameras'], featmaps=fmap)
        rgb = net(rgb_feat, ray_diff, mask, pts, ray_d)
        if rgb.shape[1] > 3:
            rgb, weights = rgb[:, 0:3], rgb[:, 3:]
            depth_map = torch.sum(weights * z, dim=-1)
        else:
            weights, depth_map = None, None
        return {'rgb': rgb, 'weights': weights, 'depth': depth_map}

    ret['outputs_coarse'] = level(model.net_coarse, featmaps[0], z_vals)
    if N_importance > 0:
        weights = ret['outputs_coarse']['weights'].clone().detach()
        z_fine = _fine_z(z_vals, weights, N_importance, inv_uniform, det)
        net = model.net_coarse if single_net else model.net_fine
        ret['outputs_fine'] = level(net, featmaps[1], z_fine)
    return ret
