"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference) on
seeded synthetic inputs.  Runs only in the build container (the reference does not travel to the GPU
box); the resulting fixtures are committed.  Usage:  python oracle/make_golden.py

Every array a fixture holds was produced by reference code:
  Projector.compute        /root/reference/ibrnet/projection.py:89-132
  IBRNet(...)/forward      /root/reference/ibrnet/mlp_network.py:152-274
  sample_along_camera_ray, sample_pdf, raw2outputs, render_rays   /root/reference/ibrnet/render_ray.py
and gradients by torch.autograd through those same functions.
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

from ibrnet.projection import Projector            # noqa: E402  (reference)
from ibrnet.mlp_network import IBRNet              # noqa: E402  (reference)
from ibrnet import render_ray as ref_rr            # noqa: E402  (reference)

from nerfool_b200.synthetic import make_scene, ray_batch_for   # noqa: E402

OUT = os.path.join(REPO, 'tests', 'golden')


def np_state(net):
    return {k: v.detach().numpy().copy() for k, v in net.state_dict().items()}


def build_nets(seed, s_c, s_f, sigma_bias):
    args = types.SimpleNamespace(anti_alias_pooling=1, local_rank=0)
    torch.manual_seed(seed)
    nc = IBRNet(args, in_feat_ch=32, n_samples=s_c)
    nf = IBRNet(args, in_feat_ch=32, n_samples=s_f)
    with torch.no_grad():
        for n in (nc, nf):
            n.out_geometry_fc[2].bias += sigma_bias
            # random-init biases are zero on most layers; perturb so bias handling is exercised
            for name, prm in n.named_parameters():
                if name.endswith('.bias'):
                    prm += 0.05 * torch.randn_like(prm)
    return nc.eval(), nf.eval()


def golden_render(name, H, W, V, R, s_c, n_imp, seed, kind, inv_uniform, white_bkgd=False, sigma_bias=0.5):
    scene = make_scene(H, W, V, seed=seed, kind=kind)
    rs = np.random.RandomState(seed + 1)
    ids = np.sort(rs.choice(H * W, R, replace=False))
    batch = ray_batch_for(scene, ids)
    nc, nf = build_nets(seed, s_c, s_c + n_imp, sigma_bias)
    fm = tuple(f.clone().requires_grad_(True) for f in scene['featmaps'])
    model = types.SimpleNamespace(net_coarse=nc, net_fine=nf)
    proj = Projector(device='cpu')

    # --- stage-level reference outputs (coarse level) ---
    pts, z = ref_rr.sample_along_camera_ray(batch['ray_o'], batch['ray_d'], batch['depth_range'], s_c,
                                            inv_uniform=inv_uniform, det=True)
    rgb_feat, ray_diff, mask = proj.compute(pts, batch['camera'], batch['src_rgbs'], batch['src_cameras'],
                                            featmaps=fm[0])
    raw_c = nc(rgb_feat, ray_diff, mask)
    # gradient of IBRNet.forward alone w.r.t. its rgb_feat input, for a fixed cotangent
    cot = torch.from_numpy(np.random.RandomState(seed + 2).randn(*raw_c.shape).astype(np.float32))
    rf = rgb_feat.detach().clone().requires_grad_(True)
    (nc(rf, ray_diff.detach(), mask.detach()) * cot).sum().backward()
    d_rgb_feat = rf.grad.clone()
    for prm in nc.parameters():
        prm.grad = None

    ret = ref_rr.render_rays(batch, model, fm, proj, N_samples=s_c, inv_uniform=inv_uniform,
                             N_importance=n_imp, det=True, white_bkgd=white_bkgd)
    gt = batch['rgb']

    def mse(o):
        m = o['mask'].float()
        return torch.sum((o['rgb'] - gt) ** 2 * m[:, None]) / (torch.sum(m) * 3 + 1e-6)
    loss = mse(ret['outputs_coarse']) + mse(ret['outputs_fine'])
    loss.backward()

    arrs = dict(
        H=H, W=W, V=V, R=R, S_c=s_c, N_imp=n_imp, inv_uniform=int(inv_uniform), white_bkgd=int(white_bkgd),
        ray_ids=ids, ray_o=batch['ray_o'].numpy(), ray_d=batch['ray_d'].numpy(),
        depth_range=batch['depth_range'].numpy(), camera=batch['camera'].numpy(),
        src_cameras=batch['src_cameras'].numpy(), src_rgbs=batch['src_rgbs'].numpy(), gt_rgb=gt.numpy(),
        feat_c=scene['featmaps'][0].numpy(), feat_f=scene['featmaps'][1].numpy(),
        pts_c=pts.numpy(), z_c=z.numpy(),
        rgb_feat_c=rgb_feat.detach().numpy(), ray_diff_c=ray_diff.numpy(), mask_c=mask.numpy(),
        raw_c=raw_c.detach().numpy(), cot_raw_c=cot.numpy(), d_rgb_feat_c=d_rgb_feat.numpy(),
        loss=np.float32(loss.item()),
        d_feat_c=fm[0].grad.numpy(), d_feat_f=fm[1].grad.numpy(),
    )
    for lvl in ('coarse', 'fine'):
        o = ret['outputs_' + lvl]
        for k in ('rgb', 'depth', 'weights', 'mask', 'alpha', 'z_vals'):
            arrs[f'{lvl}_{k}'] = o[k].detach().numpy()
    for tag, net in (('nc', nc), ('nf', nf)):
        for k, v in np_state(net).items():
            arrs[f'{tag}.{k}'] = v
    path = os.path.join(OUT, name + '.npz')
    np.savez_compressed(path, **arrs)
    print(name, 'loss', loss.item(), 'mask frac', mask.mean().item(),
          'ray-mask', ret['outputs_fine']['mask'].float().mean().item(), os.path.getsize(path) // 1024, 'KiB')


def golden_sample_pdf(seed=7, R=512, M=62, N=64):
    """Op-level vectors for sample_pdf (render_ray.py:24-70), det and non-det, flipped-inverse-depth bins
    as the inv_uniform call site builds them (:220-227).  Also stores the reference's internal cdf,
    recomputed with the reference's exact expressions, so index parity can be tested given identical cdf."""
    torch.manual_seed(seed)
    z = 1. / torch.linspace(1 / 2.0, 1 / 12.0, M + 2)[None].repeat(R, 1)
    inv = 1. / z
    bins = torch.flip(.5 * (inv[:, 1:] + inv[:, :-1]), dims=[1])
    w = torch.rand(R, M) ** 6
    w[: R // 8] = 0.0                                   # empty rays: uniform pdf from the +1e-5 floor
    w[R // 8: R // 4, 10:] = 0.0                        # long flat tails: denom < 1e-5 branch
    out = {}
    for det in (True, False):
        if det:
            u = torch.linspace(0., 1., N)[None].repeat(R, 1)
            samples = ref_rr.sample_pdf(bins, w.clone(), N, det=True)
        else:
            torch.manual_seed(seed + 1)
            samples = ref_rr.sample_pdf(bins, w.clone(), N, det=False)
            torch.manual_seed(seed + 1)
            u = torch.rand(R, N)
        ww = w.clone() + 1e-5
        pdf = ww / torch.sum(ww, dim=-1, keepdim=True)
        cdf = torch.cat([torch.zeros(R, 1), torch.cumsum(pdf, dim=-1)], dim=-1)
        above = torch.zeros_like(u, dtype=torch.long)
        for i in range(M):
            above += (u >= cdf[:, i:i + 1]).long()
        tag = 'det' if det else 'rnd'
        out.update({f'u_{tag}': u.numpy(), f'samples_{tag}': samples.numpy(), f'cdf_{tag}': cdf.numpy(),
                    f'above_{tag}': above.numpy()})
    np.savez_compressed(os.path.join(OUT, 'sample_pdf.npz'), bins=bins.numpy(), weights=w.numpy(), **out)
    print('sample_pdf ok')


def golden_feature_sizes():
    """Feature-map sizes of the reference encoder for the image sizes BASELINE.json names
    (feature_network.py:231-267) - pins nerfool_b200.synthetic.feature_map_size."""
    from ibrnet.feature_network import ResUNet
    net = ResUNet(coarse_out_ch=32, fine_out_ch=32, coarse_only=False).eval()
    rows = []
    with torch.no_grad():
        for (H, W) in ((378, 504), (30, 40), (61, 83), (96, 128), (200, 200)):
            c, f = net(torch.zeros(1, 3, H, W))
            assert c.shape == f.shape
            rows.append([H, W, c.shape[2], c.shape[3]])
    np.savez_compressed(os.path.join(OUT, 'feature_sizes.npz'), table=np.array(rows))
    print('feature sizes', rows)


if __name__ == '__main__':
    os.makedirs(OUT, exist_ok=True)
    golden_render('render_llff_v3', H=30, W=40, V=3, R=40, s_c=16, n_imp=16, seed=11, kind='llff',
                  inv_uniform=True)
    golden_render('render_synth_v5', H=36, W=36, V=5, R=24, s_c=12, n_imp=20, seed=23, kind='synthetic',
                  inv_uniform=False, white_bkgd=True)
    golden_sample_pdf()
    golden_feature_sizes()
