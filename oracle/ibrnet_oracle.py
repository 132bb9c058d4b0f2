"""CPU oracle for the NeRFool / IBRNet per-ray hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``nerfool_b200/`` imports this file; it is used by
``tests/``, by ``__graft_entry__.smoke()`` and by ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs as the checker / CPU baseline.  The product path is the CUDA library and fails loudly without it.

What it is: a functional (no ``nn.Module``) restatement, in plain fp32 PyTorch CPU ops, of the reference
algorithm on the path

    Projector.compute            /root/reference/ibrnet/projection.py:24-132
    IBRNet.forward               /root/reference/ibrnet/mlp_network.py:23-43,69-119,145-149,222-274
    sample_along_camera_ray      /root/reference/ibrnet/render_ray.py:73-116
    sample_pdf                   /root/reference/ibrnet/render_ray.py:24-70
    raw2outputs                  /root/reference/ibrnet/render_ray.py:123-170
    render_rays                  /root/reference/ibrnet/render_ray.py:173-256
    render_rays_hybrid           /root/reference/ibrnet/render_ray.py:261-390
    render_single_image          /root/reference/ibrnet/render_image.py:21-123
    img2mse (masked MSE)         /root/reference/utils.py:48-58

Parity pinning: the reference ships no tests / golden vectors, so the oracle is pinned
against outputs of the reference itself, generated in the build container by ``oracle/make_golden.py`` and
``oracle/make_golden_baseline.py`` (BASELINE-shaped scenes, render_single_image, render_rays_hybrid; both import /root/reference) and committed under ``tests/golden/``.  ``tests/test_oracle_golden.py``
checks every function here against those vectors.

One deliberate, documented deviation: the normaliser of ``sample_pdf`` (render_ray.py:36, a 62-term
``torch.sum``) is accumulated in fp64 and rounded once.  torch's CPU ``sum`` uses an ISA-dependent SIMD
tree (measured: it equals the correctly rounded sum on only ~56 % of rows on AVX512) whereas ``cumsum`` /
``cumprod`` accumulate in fp64 on CPU (measured: 100 % match), so the fp64 form is the only one that is
reproducible across hosts and on the GPU.  Effect on indices: ties within 1 ulp of a CDF entry only.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

TINY = 1e-8


# --------------------------------------------------------------------------------------------------
# camera helpers
# --------------------------------------------------------------------------------------------------
def split_camera(cam: torch.Tensor):
    """cam [..., 34] -> (H, W, K[...,4,4], c2w[...,4,4]).  Layout: projection.py:46,52-53."""
    return cam[..., 0], cam[..., 1], cam[..., 2:18].reshape(*cam.shape[:-1], 4, 4), \
        cam[..., 18:34].reshape(*cam.shape[:-1], 4, 4)


def world_to_pixel_matrices(src_cams: torch.Tensor) -> torch.Tensor:
    """P_v = K_v @ inverse(c2w_v), [V,4,4]  (projection.py:52-56)."""
    _, _, K, c2w = split_camera(src_cams)
    return K.bmm(torch.inverse(c2w))


def project_points(xyz: torch.Tensor, src_cams: torch.Tensor):
    """projection.py:42-62.  xyz [R,S,3], src_cams [V,34] -> pix [V,R,S,2], in_front [V,R,S] (bool)."""
    lead = xyz.shape[:2]
    flat = xyz.reshape(-1, 3)
    V = src_cams.shape[0]
    homog = torch.cat([flat, torch.ones_like(flat[:, :1])], dim=-1)            # [N,4]
    P = world_to_pixel_matrices(src_cams)                                      # [V,4,4]
    proj = P.bmm(homog.t()[None].repeat(V, 1, 1)).permute(0, 2, 1)             # [V,N,4]
    pix = proj[..., :2] / torch.clamp(proj[..., 2:3], min=1e-8)
    pix = torch.clamp(pix, min=-1e6, max=1e6)
    in_front = proj[..., 2] > 0
    return pix.reshape(V, *lead, 2), in_front.reshape(V, *lead)


def in_image(pix: torch.Tensor, h, w) -> torch.Tensor:
    """projection.py:24-35: closed interval [0,w-1]x[0,h-1] on un-normalised pixel coordinates."""
    x, y = pix[..., 0], pix[..., 1]
    return (x <= w - 1.) & (x >= 0) & (y <= h - 1.) & (y >= 0)


def to_grid(pix: torch.Tensor, h, w) -> torch.Tensor:
    """projection.py:37-40: 2*pix/[w-1,h-1]-1 (image size, also for the smaller feature maps)."""
    scale = torch.tensor([w - 1., h - 1.]).to(pix.device)[None, None, :]
    return 2 * pix / scale - 1.


def view_angle_features(xyz: torch.Tensor, tgt_cam: torch.Tensor, src_cams: torch.Tensor) -> torch.Tensor:
    """projection.py:64-87 -> [V,R,S,4] = (unit(a-b), a.b), a/b unit dirs point->target/source centre."""
    lead = xyz.shape[:2]
    flat = xyz.reshape(-1, 3)
    src_centres = src_cams[:, -16:].reshape(-1, 4, 4)[:, :3, 3]                # [V,3]
    V = src_centres.shape[0]
    tgt_centre = tgt_cam[-16:].reshape(-1, 4, 4).repeat(V, 1, 1)[:, :3, 3]     # [V,3]
    a = tgt_centre.unsqueeze(1) - flat.unsqueeze(0)
    a = a / (torch.norm(a, dim=-1, keepdim=True) + 1e-6)
    b = src_centres.unsqueeze(1) - flat.unsqueeze(0)
    b = b / (torch.norm(b, dim=-1, keepdim=True) + 1e-6)
    d = a - b
    d_len = torch.norm(d, dim=-1, keepdim=True)
    dot = torch.sum(a * b, dim=-1, keepdim=True)
    out = torch.cat([d / torch.clamp(d_len, min=1e-6), dot], dim=-1)
    return out.reshape(V, *lead, 4)


def projector_compute(xyz, query_camera, train_imgs, train_cameras, featmaps, detach_cameras=True):
    """projection.py:89-132 (``detach_cameras=False``: gnt/projection.py:84-132, which keeps the source cameras in the
    graph).  Same argument shapes as ``Projector.compute``:
    xyz [R,S,3]; query_camera [1,34]; train_imgs [1,V,H,W,3]; train_cameras [1,V,34]; featmaps [V,C,h',w'].
    Returns rgb_feat [R,S,V,3+C], ray_diff [R,S,V,4], mask [R,S,V,1]."""
    assert train_imgs.shape[0] == 1 and train_cameras.shape[0] == 1 and query_camera.shape[0] == 1
    cams = (train_cameras.detach() if detach_cameras else train_cameras)[0]
    imgs = train_imgs[0].permute(0, 3, 1, 2)                                   # [V,3,H,W]
    tgt = query_camera[0]
    h, w = cams[0][:2]
    pix, in_front = project_points(xyz, cams)
    grid = to_grid(pix, h, w)                                                  # [V,R,S,2]
    rgb = F.grid_sample(imgs, grid, align_corners=True).permute(2, 3, 0, 1)    # [R,S,V,3]
    feat = F.grid_sample(featmaps, grid, align_corners=True).permute(2, 3, 0, 1)
    rgb_feat = torch.cat([rgb, feat], dim=-1)
    ray_diff = view_angle_features(xyz, tgt, cams).permute(1, 2, 0, 3)
    mask = (in_image(pix, h, w) * in_front).float().permute(1, 2, 0)[..., None]
    return rgb_feat, ray_diff, mask


def bilinear_gather_explicit(src: torch.Tensor, grid: torch.Tensor) -> torch.Tensor:
    """Explicit restatement of F.grid_sample(bilinear, zeros, align_corners=True),
    used to document the tap/weight semantics the CUDA gather follows.  src [V,C,h,w], grid [V,R,S,2]
    -> [V,C,R,S]."""
    V, C, h, w = src.shape
    ix = (grid[..., 0] + 1) * ((w - 1) / 2)
    iy = (grid[..., 1] + 1) * ((h - 1) / 2)
    x0 = torch.floor(ix)
    y0 = torch.floor(iy)
    tx = ix - x0
    ty = iy - y0
    out = torch.zeros(V, C, *grid.shape[1:3], dtype=src.dtype)
    vidx = torch.arange(V)[:, None, None].expand_as(x0)
    for dy, dx, wgt in ((0, 0, (1 - ty) * (1 - tx)), (0, 1, (1 - ty) * tx),
                        (1, 0, ty * (1 - tx)), (1, 1, ty * tx)):
        xx = x0 + dx
        yy = y0 + dy
        ok = (xx >= 0) & (xx <= w - 1) & (yy >= 0) & (yy <= h - 1)
        xi = xx.clamp(0, w - 1).long()
        yi = yy.clamp(0, h - 1).long()
        tap = src[vidx, :, yi, xi]                                             # [V,R,S,C]
        out += (tap * (wgt * ok)[..., None]).permute(0, 3, 1, 2)
    return out


# --------------------------------------------------------------------------------------------------
# IBRNet aggregation network (functional; ``p`` is a state_dict-like mapping of tensors)
# --------------------------------------------------------------------------------------------------
def posenc_table(d_hid: int, n_samples: int) -> torch.Tensor:
    """mlp_network.py:210-220 -> [1,n_samples,d_hid] float32 (computed in float64 numpy, then cast)."""
    pos = np.arange(n_samples, dtype=np.float64)[:, None]
    j = np.arange(d_hid)
    ang = pos / np.power(10000, 2 * (j // 2) / d_hid)[None, :]
    tab = ang.copy()
    tab[:, 0::2] = np.sin(ang[:, 0::2])
    tab[:, 1::2] = np.cos(ang[:, 1::2])
    return torch.from_numpy(tab).float().unsqueeze(0)


def _lin(p, name, x):
    return F.linear(x, p[name + '.weight'], p.get(name + '.bias'))


def _weighted_mean_var(x, w):
    """mlp_network.py:145-149 (reduction over the view axis, dim=2)."""
    mean = torch.sum(x * w, dim=2, keepdim=True)
    var = torch.sum(w * (x - mean) ** 2, dim=2, keepdim=True)
    return mean, var


def ray_self_attention(p, x, row_mask, n_head=4, d_k=4, prefix='ray_attention'):
    """mlp_network.py:69-119 with 23-43 inlined.  x [R,S,16]; row_mask [R,S,1] float.
    NB (replicated quirk): the mask is broadcast along the *key* axis, i.e. it masks whole QUERY rows
    (masked rows get uniform attention), mlp_network.py:105-106,35-36."""
    R, S, D = x.shape
    q = F.linear(x, p[prefix + '.w_qs.weight']).view(R, S, n_head, d_k).transpose(1, 2)
    k = F.linear(x, p[prefix + '.w_ks.weight']).view(R, S, n_head, d_k).transpose(1, 2)
    v = F.linear(x, p[prefix + '.w_vs.weight']).view(R, S, n_head, d_k).transpose(1, 2)
    scores = torch.matmul(q / (d_k ** 0.5), k.transpose(2, 3))                # [R,h,S,S]
    scores = scores.masked_fill(row_mask.unsqueeze(1) == 0, -1e9)
    attn = F.softmax(scores, dim=-1)
    o = torch.matmul(attn, v).transpose(1, 2).contiguous().view(R, S, -1)
    o = F.linear(o, p[prefix + '.fc.weight']) + x
    return F.layer_norm(o, (D,), p[prefix + '.layer_norm.weight'], p[prefix + '.layer_norm.bias'], eps=1e-6)


def ibrnet_forward(p, pos_encoding, rgb_feat, ray_diff, mask, anti_alias_pooling=True, want=None):
    """mlp_network.py:222-274.  rgb_feat [R,S,V,35], ray_diff [R,S,V,4], mask [R,S,V,1] -> raw [R,S,4].
    ``want``: optional dict that receives named intermediates (for stage-level parity debugging)."""
    V = rgb_feat.shape[2]
    dir_feat = F.elu(_lin(p, 'ray_dir_fc.2', F.elu(_lin(p, 'ray_dir_fc.0', ray_diff))))
    rgb_in = rgb_feat[..., :3]
    x0 = rgb_feat + dir_feat
    if anti_alias_pooling:
        dot = ray_diff[..., 3:4]
        e = torch.exp(torch.abs(p['s']) * (dot - 1))
        w = (e - torch.min(e, dim=2, keepdim=True)[0]) * mask
        w = w / (torch.sum(w, dim=2, keepdim=True) + TINY)
    else:
        w = mask / (torch.sum(mask, dim=2, keepdim=True) + TINY)
    mean0, var0 = _weighted_mean_var(x0, w)
    g = torch.cat([mean0, var0], dim=-1)
    x = torch.cat([g.expand(-1, -1, V, -1), x0], dim=-1)
    x = F.elu(_lin(p, 'base_fc.2', F.elu(_lin(p, 'base_fc.0', x))))
    xv = F.elu(_lin(p, 'vis_fc.2', F.elu(_lin(p, 'vis_fc.0', x * w))))
    x_res, vis = xv[..., :-1], xv[..., -1:]
    vis = torch.sigmoid(vis) * mask
    x = x + x_res
    vis = torch.sigmoid(_lin(p, 'vis_fc2.2', F.elu(_lin(p, 'vis_fc2.0', x * vis)))) * mask
    w2 = vis / (torch.sum(vis, dim=2, keepdim=True) + TINY)
    mean2, var2 = _weighted_mean_var(x, w2)
    gf_in = torch.cat([mean2.squeeze(2), var2.squeeze(2), w2.mean(dim=2)], dim=-1)   # [R,S,65]
    gf = F.elu(_lin(p, 'geometry_fc.2', F.elu(_lin(p, 'geometry_fc.0', gf_in))))
    n_valid = torch.sum(mask, dim=2)                                          # [R,S,1]
    gf = gf + pos_encoding
    gf = ray_self_attention(p, gf, (n_valid > 1).float())
    sigma = F.relu(_lin(p, 'out_geometry_fc.2', F.elu(_lin(p, 'out_geometry_fc.0', gf))))
    sigma = sigma.masked_fill(n_valid < 1, 0.)
    c = torch.cat([x, vis, ray_diff], dim=-1)
    c = _lin(p, 'rgb_fc.4', F.elu(_lin(p, 'rgb_fc.2', F.elu(_lin(p, 'rgb_fc.0', c)))))
    c = c.masked_fill(mask == 0, -1e9)
    blend = F.softmax(c, dim=2)
    rgb_out = torch.sum(rgb_in * blend, dim=2)
    if want is not None:
        want.update(dir_feat=dir_feat, w=w, mean0=mean0, var0=var0, x2=x, vis2=vis, w2=w2,
                    gf_in=gf_in, n_valid=n_valid, gf=gf, blend=blend, rgb_out=rgb_out)
    return torch.cat([rgb_out, sigma], dim=-1)


# --------------------------------------------------------------------------------------------------
# ray-level helpers
# --------------------------------------------------------------------------------------------------
def coarse_depths(ray_o, ray_d, depth_range, n_samples, inv_uniform=False, det=False, t_rand=None):
    """render_ray.py:73-116.  Returns pts [R,S,3], z_vals [R,S].  The per-index python loop of the
    reference (start + i*step with an int i) is kept so every z value is bit-identical."""
    near_v, far_v = depth_range[0, 0], depth_range[0, 1]
    assert near_v > 0 and far_v > 0 and far_v > near_v
    near = near_v * torch.ones_like(ray_d[..., 0])
    far = far_v * torch.ones_like(ray_d[..., 0])
    if inv_uniform:
        start = 1. / near
        step = (1. / far - start) / (n_samples - 1)
        z = 1. / torch.stack([start + i * step for i in range(n_samples)], dim=1)
    else:
        start = near
        step = (far - near) / (n_samples - 1)
        z = torch.stack([start + i * step for i in range(n_samples)], dim=1)
    if not det:
        mids = .5 * (z[:, 1:] + z[:, :-1])
        upper = torch.cat([mids, z[:, -1:]], dim=-1)
        lower = torch.cat([z[:, :1], mids], dim=-1)
        if t_rand is None:
            t_rand = torch.rand_like(z)
        z = lower + (upper - lower) * t_rand
    pts = z.unsqueeze(2) * ray_d.unsqueeze(1) + ray_o.unsqueeze(1)
    return pts, z


def cdf_from_weights(weights: torch.Tensor) -> torch.Tensor:
    """render_ray.py:33-38 -> cdf [R,M+1].  weights must already be a private copy (the reference adds
    1e-5 in place).  Normaliser accumulated in fp64 (see module docstring)."""
    w = weights + 1e-5
    total = w.double().sum(dim=-1, keepdim=True).float()
    pdf = w / total
    cdf = torch.cumsum(pdf, dim=-1)        # CPU cumsum accumulates fp32 inputs in fp64
    return torch.cat([torch.zeros_like(cdf[:, :1]), cdf], dim=-1)


def invert_cdf(bins, cdf, u):
    """render_ray.py:47-68.  bins [R,M+1], cdf [R,M+1], u [R,N] -> samples [R,N], above_inds [R,N]
    (int64).  ``above`` counts the first M cdf entries that are <= u."""
    M = cdf.shape[1] - 1
    above = (u.unsqueeze(-1) >= cdf[:, None, :M]).sum(dim=-1)                 # [R,N] int64
    below = torch.clamp(above - 1, min=0)
    c_lo = torch.gather(cdf, 1, below)
    c_hi = torch.gather(cdf, 1, above)
    b_lo = torch.gather(bins, 1, below)
    b_hi = torch.gather(bins, 1, above)
    denom = c_hi - c_lo
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    t = (u - c_lo) / denom
    return b_lo + t * (b_hi - b_lo), above


def sample_pdf(bins, weights, n_samples, det=False, u=None, return_inds=False):
    """render_ray.py:24-70."""
    cdf = cdf_from_weights(weights)
    if u is None:
        if det:
            u = torch.linspace(0., 1., n_samples).to(bins.device).unsqueeze(0).repeat(bins.shape[0], 1)
        else:
            u = torch.rand(bins.shape[0], n_samples, device=bins.device)
    samples, above = invert_cdf(bins, cdf, u)
    return (samples, above) if return_inds else samples


def fine_depths(z_coarse, weights_coarse, n_importance, inv_uniform=False, det=False, u=None):
    """render_ray.py:216-238 -> sorted z_vals [R, S+N_importance] (weights are treated as constants)."""
    w = weights_coarse.clone().detach()[:, 1:-1]
    if inv_uniform:
        inv_z = 1. / z_coarse
        inv_mid = .5 * (inv_z[:, 1:] + inv_z[:, :-1])
        inv_s = sample_pdf(torch.flip(inv_mid, dims=[1]), torch.flip(w, dims=[1]), n_importance, det=det, u=u)
        z_new = 1. / inv_s
    else:
        mid = .5 * (z_coarse[:, 1:] + z_coarse[:, :-1])
        z_new = sample_pdf(mid, w, n_importance, det=det, u=u)
    z_all, _ = torch.sort(torch.cat((z_coarse, z_new), dim=-1), dim=-1)
    return z_all


def composite(raw, z_vals, pixel_mask, white_bkgd=False):
    """render_ray.py:123-170 (geo_noise=None path; the interval is not used, :136-139)."""
    rgb, sigma = raw[:, :, :3], raw[:, :, 3]
    alpha = 1. - torch.exp(-sigma)
    T = torch.cumprod(1. - alpha + 1e-10, dim=-1)[:, :-1]
    T = torch.cat((torch.ones_like(T[:, :1]), T), dim=-1)
    weights = alpha * T
    rgb_map = torch.sum(weights.unsqueeze(2) * rgb, dim=1)
    if white_bkgd:
        rgb_map = rgb_map + (1. - torch.sum(weights, dim=-1, keepdim=True))
    ray_mask = pixel_mask.float().sum(dim=1) > 8
    depth_map = torch.sum(weights * z_vals, dim=-1)
    return OrderedDict([('rgb', rgb_map), ('depth', depth_map), ('weights', weights),
                        ('mask', ray_mask), ('alpha', alpha), ('z_vals', z_vals)])


def render_rays(ray_batch, params_coarse, params_fine, featmaps, n_samples, inv_uniform=False,
                n_importance=0, det=False, white_bkgd=False, anti_alias_pooling=True, src_ray_batch=None,
                u=None, t_rand=None, fine_z=None):
    """render_ray.py:173-256.  ``params_*`` are state_dict-like mappings that also hold 'pos_encoding'.
    ``fine_z`` (test hook, not in the reference): evaluate the fine level at these depths instead of sampling them -- used
    to compare two precisions of the same arithmetic at IDENTICAL sample positions."""
    src = ray_batch if src_ray_batch is None else src_ray_batch
    pts, z = coarse_depths(ray_batch['ray_o'], ray_batch['ray_d'], ray_batch['depth_range'],
                           n_samples, inv_uniform=inv_uniform, det=det, t_rand=t_rand)
    rgb_feat, ray_diff, mask = projector_compute(pts, ray_batch['camera'], src['src_rgbs'],
                                                 src['src_cameras'], featmaps[0])
    pixel_mask = mask[..., 0].sum(dim=2) > 1
    raw = ibrnet_forward(params_coarse, params_coarse['pos_encoding'], rgb_feat, ray_diff, mask,
                         anti_alias_pooling)
    out = {'outputs_coarse': composite(raw, z, pixel_mask, white_bkgd), 'outputs_fine': None}
    if n_importance > 0:
        z = fine_depths(z, out['outputs_coarse']['weights'], n_importance, inv_uniform, det, u=u) if fine_z is None else fine_z
        pts = z.unsqueeze(2) * ray_batch['ray_d'].unsqueeze(1) + ray_batch['ray_o'].unsqueeze(1)
        rgb_feat, ray_diff, mask = projector_compute(pts, ray_batch['camera'], src['src_rgbs'],
                                                     src['src_cameras'], featmaps[1])
        pixel_mask = mask[..., 0].sum(dim=2) > 1
        raw = ibrnet_forward(params_fine, params_fine['pos_encoding'], rgb_feat, ray_diff, mask,
                             anti_alias_pooling)
        out['outputs_fine'] = composite(raw, z, pixel_mask, white_bkgd)
    return out


def render_rays_hybrid(ray_batch, params_coarse, params_fine, featmaps, featmaps_clean, n_samples, use_clean_color,
                       use_clean_density, inv_uniform=False, n_importance=0, det=False, white_bkgd=False,
                       anti_alias_pooling=True, src_ray_batch=None):
    """render_ray.py:261-390: both feature-map sets at the same points; colour and density each from the clean or the
    adversarial pass; composited with the ADVERSARIAL pass's pixel mask (:320-321, :386-387)."""
    src = ray_batch if src_ray_batch is None else src_ray_batch

    def level(params, fm_adv, fm_clean, z):
        pts = z.unsqueeze(2) * ray_batch['ray_d'].unsqueeze(1) + ray_batch['ray_o'].unsqueeze(1)
        raws, pixel_mask = [], None
        for fm in (fm_adv, fm_clean):
            rgb_feat, ray_diff, mask = projector_compute(pts, ray_batch['camera'], src['src_rgbs'], src['src_cameras'], fm)
            if pixel_mask is None:
                pixel_mask = mask[..., 0].sum(dim=2) > 1
            raws.append(ibrnet_forward(params, params['pos_encoding'], rgb_feat, ray_diff, mask, anti_alias_pooling))
        color = raws[1][:, :, :3] if use_clean_color else raws[0][:, :, :3]
        sigma = raws[1][:, :, 3:4] if use_clean_density else raws[0][:, :, 3:4]
        return composite(torch.cat([color, sigma], dim=2), z, pixel_mask, white_bkgd)

    _, z = coarse_depths(ray_batch['ray_o'], ray_batch['ray_d'], ray_batch['depth_range'], n_samples,
                         inv_uniform=inv_uniform, det=det)
    out = {'outputs_coarse': level(params_coarse, featmaps[0], featmaps_clean[0], z), 'outputs_fine': None}
    if n_importance > 0:
        z = fine_depths(z, out['outputs_coarse']['weights'], n_importance, inv_uniform, det)
        out['outputs_fine'] = level(params_fine, featmaps[1], featmaps_clean[1], z)
    return out


def render_single_image(H, W, ray_batch, params_coarse, params_fine, featmaps, chunk_size, n_samples, inv_uniform=False,
                        n_importance=0, det=False, white_bkgd=False):
    """render_image.py:21-123 (render_stride 1, plain render_rays branch): chunk loop over the rays of a view, outputs
    concatenated and reshaped to [H, W, ...] (squeezed), masked-out pixels of the COARSE image painted white (:109)."""
    shared = ('camera', 'depth_range', 'src_rgbs', 'src_cameras')
    acc = {'outputs_coarse': {}, 'outputs_fine': {}}
    n_rays = ray_batch['ray_o'].shape[0]
    for i in range(0, n_rays, chunk_size):
        chunk = {k: (v if k in shared or v is None else v[i:i + chunk_size]) for k, v in ray_batch.items()}
        ret = render_rays(chunk, params_coarse, params_fine, featmaps, n_samples, inv_uniform, n_importance, det, white_bkgd)
        for lvl in acc:
            if ret[lvl] is None:
                acc[lvl] = None
                continue
            for k, v in ret[lvl].items():
                acc[lvl].setdefault(k, []).append(v)
    out = {}
    for lvl, d in acc.items():
        out[lvl] = None if d is None else OrderedDict((k, torch.cat(v, dim=0).reshape(H, W, -1).squeeze()) for k, v in d.items())
    out['outputs_coarse']['rgb'][out['outputs_coarse']['mask'] == 0] = 1.
    return out


def fine_depth_report(z_ours, z_ref, n_coarse_sorted=None):
    """How two fine-depth tensors [R, S] differ, for the end-to-end z_vals parity statement (render_ray.py:216-238): the
    importance samples are CONTINUOUS piecewise-linear functions of the coarse weights (inverse CDF), so a relative
    weight error e moves a sample by <= e x (bin width) -- except that merge-sort ranks swap when two depths cross, which
    is harmless (same multiset up to e).  Returns max |dz|, max |dz| / bin-scale, and the number of rays whose sorted
    sequences differ by more than `n ulp` anywhere."""
    dz = (z_ours.double() - z_ref.double()).abs()
    scale = (z_ref[:, -1] - z_ref[:, 0]).double().clamp_min(1e-12).unsqueeze(1)
    return {'max_abs': dz.max().item(), 'max_rel_range': (dz / scale).max().item(),
            'n_exact': int((dz == 0).sum().item()), 'n': dz.numel()}


def masked_mse(x, y, mask=None):
    """utils.py:48-58."""
    if mask is None:
        return torch.mean((x - y) * (x - y))
    return torch.sum((x - y) * (x - y) * mask.unsqueeze(-1)) / (torch.sum(mask) * x.shape[-1] + 1e-6)


def attack_loss(out, gt_rgb):
    """criterion.py:23-33 summed over coarse+fine as in eval_adv.py:306-310."""
    loss = masked_mse(out['outputs_coarse']['rgb'], gt_rgb, out['outputs_coarse']['mask'].float())
    if out['outputs_fine'] is not None:
        loss = loss + masked_mse(out['outputs_fine']['rgb'], gt_rgb, out['outputs_fine']['mask'].float())
    return loss


# --------------------------------------------------------------------------------------------------
# parameter construction (same shapes / names / init scheme as mlp_network.py:153-208)
# --------------------------------------------------------------------------------------------------
IBRNET_PARAM_SHAPES = OrderedDict([
    ('s', ()),
    ('ray_dir_fc.0.weight', (16, 4)), ('ray_dir_fc.0.bias', (16,)),
    ('ray_dir_fc.2.weight', (35, 16)), ('ray_dir_fc.2.bias', (35,)),
    ('base_fc.0.weight', (64, 105)), ('base_fc.0.bias', (64,)),
    ('base_fc.2.weight', (32, 64)), ('base_fc.2.bias', (32,)),
    ('vis_fc.0.weight', (32, 32)), ('vis_fc.0.bias', (32,)),
    ('vis_fc.2.weight', (33, 32)), ('vis_fc.2.bias', (33,)),
    ('vis_fc2.0.weight', (32, 32)), ('vis_fc2.0.bias', (32,)),
    ('vis_fc2.2.weight', (1, 32)), ('vis_fc2.2.bias', (1,)),
    ('geometry_fc.0.weight', (64, 65)), ('geometry_fc.0.bias', (64,)),
    ('geometry_fc.2.weight', (16, 64)), ('geometry_fc.2.bias', (16,)),
    ('ray_attention.w_qs.weight', (16, 16)), ('ray_attention.w_ks.weight', (16, 16)),
    ('ray_attention.w_vs.weight', (16, 16)), ('ray_attention.fc.weight', (16, 16)),
    ('ray_attention.layer_norm.weight', (16,)), ('ray_attention.layer_norm.bias', (16,)),
    ('out_geometry_fc.0.weight', (16, 16)), ('out_geometry_fc.0.bias', (16,)),
    ('out_geometry_fc.2.weight', (1, 16)), ('out_geometry_fc.2.bias', (1,)),
    ('rgb_fc.0.weight', (16, 37)), ('rgb_fc.0.bias', (16,)),
    ('rgb_fc.2.weight', (8, 16)), ('rgb_fc.2.bias', (8,)),
    ('rgb_fc.4.weight', (1, 8)), ('rgb_fc.4.bias', (1,)),
])


def random_ibrnet_params(n_samples: int, seed: int, sigma_bias: float = 0.0):
    """Random-init parameters with the reference's shapes.  Kaiming-normal + zero bias where the reference
    uses ``weights_init`` (mlp_network.py:137-141,204-208), U(+-1/sqrt(fan_in)) elsewhere.  Not
    bit-identical to ``IBRNet.__init__`` under the same seed (golden tests load the reference's own
    state_dict instead); used for GPU-vs-oracle parity at sizes where no fixture is committed."""
    g = torch.Generator().manual_seed(seed)
    p = OrderedDict()
    kaiming = ('base_fc', 'vis_fc', 'vis_fc2', 'geometry_fc', 'rgb_fc')
    for name, shape in IBRNET_PARAM_SHAPES.items():
        if name == 's':
            p[name] = torch.tensor(0.2)
        elif name.endswith('layer_norm.weight'):
            p[name] = torch.ones(shape)
        elif name.endswith('layer_norm.bias'):
            p[name] = torch.zeros(shape)
        elif name.endswith('.weight'):
            fan_in = shape[1]
            if name.split('.')[0] in kaiming:
                p[name] = torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_in)
            else:
                bound = 1.0 / math.sqrt(fan_in)
                p[name] = (torch.rand(shape, generator=g) * 2 - 1) * bound
        else:
            mod = name.split('.')[0]
            if mod in kaiming:
                p[name] = torch.zeros(shape)
            else:
                fan_in = IBRNET_PARAM_SHAPES[name.replace('.bias', '.weight')][1]
                bound = 1.0 / math.sqrt(fan_in)
                p[name] = (torch.rand(shape, generator=g) * 2 - 1) * bound
    if sigma_bias:
        p['out_geometry_fc.2.bias'] = p['out_geometry_fc.2.bias'] + sigma_bias
    p['pos_encoding'] = posenc_table(16, n_samples)
    return p
