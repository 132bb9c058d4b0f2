"""Round-2 fixtures: the UNMODIFIED reference (imported from /root/reference) on BASELINE-shaped scenes, and its own
``render_single_image`` / ``render_rays_hybrid``.  Build container only; the fixtures are committed.

    python oracle/make_golden_baseline.py

  base_llff_v4.npz     378x504, V=4,  64+64   (BASELINE configs[0]/[1]),  192 rays: render_rays + autograd to the feature maps
  base_llff_v10.npz    378x504, V=10, 64+64   (BASELINE.json metric / configs[2]),  128 rays
  base_synth_v10.npz   200x200, V=10, 64+128  (configs[3] sampling, reduced image),  96 rays
  render_image.npz     ibrnet/render_image.py:21-123 on a 40x56 view (V=4, 24+24, chunk 500) -- every output map
  hybrid.npz           ibrnet/render_ray.py:261-390 for the three (use_clean_color, use_clean_density) settings
  forward_warp.npz     eval/ibrnet/eval_adv.py:97-197 (three modes, with and without zero depths) + depth-smooth / depth-var losses

The BASELINE-shaped scenes are too large to commit (9 MB of source images + 12.6 MB of feature maps at V=4), so those
fixtures hold the scene SEED plus a sha256 of every regenerated input tensor: the tests rebuild the inputs with
``nerfool_b200.synthetic.make_scene`` and check the digests before comparing against the reference's outputs."""
import hashlib
import os
import sys
import types

import numpy as np
import torch

sys.dont_write_bytecode = True
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(1, '/root/reference')

from ibrnet.projection import Projector            # noqa: E402  (reference)
from ibrnet.mlp_network import IBRNet              # noqa: E402  (reference)
from ibrnet import render_ray as ref_rr            # noqa: E402  (reference)
from ibrnet import render_image as ref_ri          # noqa: E402  (reference)

from nerfool_b200.synthetic import make_scene, ray_batch_for, rays_for_view   # noqa: E402

OUT = os.path.join(REPO, 'tests', 'golden')
GRAD_TEXELS = 3000


def digest(t):
    return hashlib.sha256(np.ascontiguousarray(t.detach().cpu().numpy()).tobytes()).hexdigest()


def scene_digests(scene):
    return {'sha_src_rgbs': digest(scene['src_rgbs']), 'sha_feat_c': digest(scene['featmaps'][0]),
            'sha_feat_f': digest(scene['featmaps'][1]), 'sha_camera': digest(scene['camera']),
            'sha_src_cameras': digest(scene['src_cameras']), 'sha_rgb': digest(scene['rgb'])}


def build_nets(seed, s_c, s_f, sigma_bias):
    args = types.SimpleNamespace(anti_alias_pooling=1, local_rank=0)
    torch.manual_seed(seed)
    nc = IBRNet(args, in_feat_ch=32, n_samples=s_c)
    nf = IBRNet(args, in_feat_ch=32, n_samples=s_f)
    with torch.no_grad():
        for n in (nc, nf):
            n.out_geometry_fc[2].bias += sigma_bias
            for name, prm in n.named_parameters():
                if name.endswith('.bias'):
                    prm += 0.05 * torch.randn_like(prm)
    return nc.eval(), nf.eval()


def state_arrays(nc, nf):
    out = {}
    for tag, net in (('nc', nc), ('nf', nf)):
        for k, v in net.state_dict().items():
            out[f'{tag}.{k}'] = v.detach().numpy().copy()
    return out


def golden_baseline(name, H, W, V, R, s_c, n_imp, seed, kind, sigma_bias=0.4):
    scene = make_scene(H, W, V, seed=seed, kind=kind)
    ids = np.sort(np.random.RandomState(seed + 1).choice(H * W, R, replace=False))
    batch = ray_batch_for(scene, ids)
    nc, nf = build_nets(seed, s_c, s_c + n_imp, sigma_bias)
    fm = tuple(f.clone().requires_grad_(True) for f in scene['featmaps'])
    model = types.SimpleNamespace(net_coarse=nc, net_fine=nf)
    ret = ref_rr.render_rays(batch, model, fm, Projector(device='cpu'), N_samples=s_c, inv_uniform=True,
                             N_importance=n_imp, det=True, white_bkgd=False)
    gt = batch['rgb']

    def mse(o):
        m = o['mask'].float()
        return torch.sum((o['rgb'] - gt) ** 2 * m[:, None]) / (torch.sum(m) * 3 + 1e-6)
    lc, lf = mse(ret['outputs_coarse']), mse(ret['outputs_fine'])
    (lc + lf).backward()
    # the feature-map gradients are sparse (R rays) but still megabytes: keep a seeded random SAMPLE of at most GRAD_TEXELS
    # non-zero texels per level (all 32 channels of each) plus the norm of the full gradient
    arrs = dict(H=H, W=W, V=V, R=R, S_c=s_c, N_imp=n_imp, seed=seed, kind=kind, ray_ids=ids, loss=np.float32((lc + lf).item()),
                loss_coarse=np.float32(lc.item()), loss_fine=np.float32(lf.item()))
    arrs.update({k: np.array(v) for k, v in scene_digests(scene).items()})
    for tag, g in (('c', fm[0].grad), ('f', fm[1].grad)):
        g = g.permute(0, 2, 3, 1).reshape(-1, 32)                     # [V*h*w, 32]
        nz = torch.nonzero(g.abs().sum(dim=1) > 0)[:, 0]
        arrs[f'd_feat_{tag}_norm'] = np.float64(g.double().norm().item())
        arrs[f'd_feat_{tag}_nnz_texels'] = np.int64(nz.numel())
        if nz.numel() > GRAD_TEXELS:
            pick = np.sort(np.random.RandomState(seed + 7).choice(nz.numel(), GRAD_TEXELS, replace=False))
            nz = nz[torch.from_numpy(pick)]
        arrs[f'd_feat_{tag}_idx'] = nz.numpy().astype(np.int32)
        arrs[f'd_feat_{tag}_val'] = g[nz].numpy()
        arrs[f'd_feat_{tag}_shape'] = np.array(fm[0].shape)
    for lvl in ('coarse', 'fine'):
        o = ret['outputs_' + lvl]
        for k in ('rgb', 'depth', 'weights', 'mask', 'alpha', 'z_vals'):
            arrs[f'{lvl}_{k}'] = o[k].detach().numpy()
    arrs.update(state_arrays(nc, nf))
    path = os.path.join(OUT, name + '.npz')
    np.savez_compressed(path, **arrs)
    print(name, 'loss', (lc + lf).item(), 'ray-mask', ret['outputs_fine']['mask'].float().mean().item(),
          os.path.getsize(path) // 1024, 'KiB')


def small_scene_arrays(scene):
    return dict(camera=scene['camera'].numpy(), src_cameras=scene['src_cameras'].numpy(), src_rgbs=scene['src_rgbs'].numpy(),
                depth_range=scene['depth_range'].numpy(), feat_c=scene['featmaps'][0].numpy(), feat_f=scene['featmaps'][1].numpy())


def golden_render_image(H=40, W=56, V=4, s_c=24, n_imp=24, seed=5):
    """render_single_image (render_image.py:21-123) with the reference's chunk loop, white-painted masked pixels."""
    scene = make_scene(H, W, V, seed=seed, kind='llff')
    o, d = rays_for_view(scene['camera'][0], H, W)
    ray_batch = {'ray_o': o, 'ray_d': d, 'depth_range': scene['depth_range'], 'camera': scene['camera'][:1],
                 'rgb': scene['rgb'][0], 'src_rgbs': scene['src_rgbs'], 'src_cameras': scene['src_cameras'],
                 'src_depths': None, 'depth': None, 'depth_full': None}       # the keys RaySamplerSingleImage.get_all returns
    nc, nf = build_nets(seed, s_c, s_c + n_imp, 0.4)
    model = types.SimpleNamespace(net_coarse=nc, net_fine=nf)
    sampler = types.SimpleNamespace(H=H, W=W)
    with torch.no_grad():
        ret = ref_ri.render_single_image(sampler, ray_batch, model, Projector(device='cpu'), 500, s_c, inv_uniform=True,
                                         N_importance=n_imp, det=True, white_bkgd=False, render_stride=1,
                                         featmaps=scene['featmaps'])
    arrs = dict(H=H, W=W, V=V, S_c=s_c, N_imp=n_imp, **small_scene_arrays(scene))
    for lvl in ('coarse', 'fine'):
        for k, v in ret['outputs_' + lvl].items():
            arrs[f'{lvl}_{k}'] = v.numpy()
    arrs.update(state_arrays(nc, nf))
    path = os.path.join(OUT, 'render_image.npz')
    np.savez_compressed(path, **arrs)
    print('render_image', {k: tuple(v.shape) for k, v in ret['outputs_fine'].items()}, os.path.getsize(path) // 1024, 'KiB')


def golden_hybrid(H=40, W=56, V=4, R=60, s_c=24, n_imp=24, seed=6):
    """render_rays_hybrid (render_ray.py:261-390): adversarial + clean feature maps, colour / density taken from either."""
    scene = make_scene(H, W, V, seed=seed, kind='llff')
    ids = np.sort(np.random.RandomState(seed).choice(H * W, R, replace=False))
    batch = ray_batch_for(scene, ids)
    nc, nf = build_nets(seed, s_c, s_c + n_imp, 0.4)
    model = types.SimpleNamespace(net_coarse=nc, net_fine=nf)
    g = torch.Generator().manual_seed(seed)
    fm_adv = tuple(f + 0.3 * torch.randn(f.shape, generator=g) for f in scene['featmaps'])
    arrs = dict(H=H, W=W, V=V, S_c=s_c, N_imp=n_imp, ray_ids=ids, ray_o=batch['ray_o'].numpy(), ray_d=batch['ray_d'].numpy(),
                adv_c=fm_adv[0].numpy(), adv_f=fm_adv[1].numpy(), **small_scene_arrays(scene))
    for cc, cd in ((1, 0), (0, 1), (1, 1)):
        a = types.SimpleNamespace(use_clean_color=bool(cc), use_clean_density=bool(cd))
        with torch.no_grad():
            ret = ref_rr.render_rays_hybrid(batch, model, fm_adv, Projector(device='cpu'), s_c, inv_uniform=True,
                                            N_importance=n_imp, det=True, white_bkgd=False, args=a,
                                            featmaps_clean=scene['featmaps'])
        for lvl in ('coarse', 'fine'):
            for k in ('rgb', 'depth', 'weights', 'mask', 'z_vals'):
                arrs[f'c{cc}d{cd}_{lvl}_{k}'] = ret['outputs_' + lvl][k].numpy()
    arrs.update(state_arrays(nc, nf))
    path = os.path.join(OUT, 'hybrid.npz')
    np.savez_compressed(path, **arrs)
    print('hybrid', os.path.getsize(path) // 1024, 'KiB')


def golden_forward_warp(H=40, W=56, seed=9, n_sel=300):
    """forward_warp (eval/ibrnet/eval_adv.py:97-197), the reference's Python z-buffer loop, in its three modes (src2tar with
    the selected-ray filter, tar2src over the selected rays, full image) and on a depth map holding zeros (the "empty" marker
    doubles as a value).  eval_adv.py is imported unmodified through tests/ref_harness.py (stub modules for absent packages)."""
    sys.path.insert(0, os.path.join(REPO, 'tests'))
    import ref_harness
    ref_harness.setup_paths('ibrnet')
    import eval_adv
    scene = make_scene(H, W, 2, seed=seed, kind='llff')
    cams = scene['src_cameras'][0]
    K = [c[2:18].reshape(4, 4)[:3, :3].clone() for c in cams]
    E = [c[18:34].reshape(4, 4).clone() for c in cams]
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing='ij')
    depth = 2.5 + 0.8 * torch.sin(xx / 7.0) * torch.cos(yy / 5.0) + 0.6 * (xx > W / 2).float()      # smooth + an occluding step
    depth_z = depth.clone()
    depth_z[torch.rand(H, W, generator=g) < 0.03] = 0.0
    rgb = torch.rand(H, W, 3, generator=g)
    sel = np.sort(np.random.RandomState(seed).choice(H * W, n_sel, replace=False))
    arrs = dict(H=H, W=W, depth=depth.numpy(), depth_zeros=depth_z.numpy(), rgb=rgb.numpy(), sel=sel,
                K0=K[0].numpy(), K1=K[1].numpy(), E0=E[0].numpy(), E1=E[1].numpy())
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        for tag, d in (('pos', depth), ('zero', depth_z)):
            for mode, kw in (('s2t', dict(src2tar=True)), ('t2s', dict(src2tar=False)), ('full', dict(derive_full_image=True))):
                out = eval_adv.forward_warp(sel, rgb, d[None], K[0], E[0], K[1], E[1], **kw)
                arrs[f'{tag}_{mode}_new'] = out[0].numpy()
                arrs[f'{tag}_{mode}_new_depth'] = out[1].numpy()
                arrs[f'{tag}_{mode}_rgb_proj'] = out[2].numpy()
                arrs[f'{tag}_{mode}_depth_proj'] = out[3].numpy()
                if mode == 't2s':
                    arrs[f'{tag}_{mode}_inds_new'] = np.asarray(out[4], dtype=np.int64)
    # depth-smooth / depth-variance losses on fixed inputs (eval_adv.py:32-48, train.py:329-340)
    import train as ref_train
    dm = torch.rand(4 * 8 * 8, generator=g) * 3 + 2
    wts = torch.rand(50, 24, generator=g) ** 3
    wts[:3] = 0                                            # zero total weight -> NaN rays, dropped
    zs = torch.sort(torch.rand(50, 24, generator=g) * 10 + 2, dim=1)[0]
    dep = (wts * zs).sum(1)
    arrs.update(ds_depth=dm.numpy(), ds_l2=eval_adv.calc_depth_smooth_loss({'depth': dm}, 8).numpy(),
                ds_l1=eval_adv.calc_depth_smooth_loss({'depth': dm}, 8, 'l1').numpy(),
                dv_weights=wts.numpy(), dv_z=zs.numpy(), dv_depth=dep.numpy(),
                dv=ref_train.calc_depth_var({'depth': dep, 'weights': wts, 'z_vals': zs}).numpy())
    path = os.path.join(OUT, 'forward_warp.npz')
    np.savez_compressed(path, **arrs)
    print('forward_warp', 'filled px (s2t/t2s/full):', int((arrs['pos_s2t_new_depth'] > 0).sum()), int((arrs['pos_t2s_new_depth'] > 0).sum()),
          int((arrs['pos_full_new_depth'] > 0).sum()), os.path.getsize(path) // 1024, 'KiB')


if __name__ == '__main__':
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == 'warp':
        golden_forward_warp()
        sys.exit(0)
    torch.set_num_threads(os.cpu_count() or 1)
    golden_baseline('base_llff_v4', 378, 504, 4, 192, 64, 64, seed=41, kind='llff')
    golden_baseline('base_llff_v10', 378, 504, 10, 128, 64, 64, seed=42, kind='llff')
    golden_baseline('base_synth_v10', 200, 200, 10, 96, 64, 128, seed=43, kind='synthetic')
    golden_render_image()
    golden_hybrid()
    golden_forward_warp()
