"""Stage-by-stage CUDA-vs-oracle error table (run on the GPU box; writes gpurun_out/diag.json).
Not a pytest file: it never asserts, it reports, so that one GPU call shows every stage at once."""
import json
import os
import sys
import time
import types

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, 'tests'))

from oracle import ibrnet_oracle as O                      # noqa: E402
from nerfool_b200 import ops, _lib                          # noqa: E402
from nerfool_b200.mlp_network import IBRNet, pack_params   # noqa: E402
from nerfool_b200.projection import Projector              # noqa: E402
from nerfool_b200 import render_ray as RR                  # noqa: E402
from nerfool_b200.synthetic import make_scene, ray_batch_for  # noqa: E402

dev = torch.device('cuda:0')
out = {}


def rec(name, a, b, exact=False):
    a = torch.as_tensor(a).detach().cpu().double()
    b = torch.as_tensor(b).detach().cpu().double()
    d = (a - b).abs()
    rel = (a - b).norm() / (b.norm() + 1e-30)
    e = {'maxabs': d.max().item() if d.numel() else 0.0, 'rel': rel.item(), 'ref_absmax': b.abs().max().item() if b.numel() else 0.0,
         'n': a.numel(), 'nan': int(torch.isnan(a).sum())}
    if exact:
        e['mismatch'] = int((a != b).sum())
    out[name] = e
    print(f'{name:44s} maxabs {e["maxabs"]:.3e} rel {e["rel"]:.3e} ref|max| {e["ref_absmax"]:.3e} nan {e["nan"]}'
          + (f' mismatches {e["mismatch"]}/{e["n"]}' if exact else ''), flush=True)


def make_net(p, S):
    args = types.SimpleNamespace(anti_alias_pooling=1)
    net = IBRNet(args, 32, S)
    sd = {k: v.clone() for k, v in p.items()}
    net.load_state_dict(sd)
    return net.to(dev).eval()


def run(V, R, S_c, n_imp, H=378, W=504, kind='llff', inv_uniform=True, seed=0, tag=''):
    tag = tag or f'V{V}_R{R}_S{S_c}+{n_imp}'
    print(f'=== {tag} ===', flush=True)
    scene = make_scene(H, W, V, seed=seed, kind=kind)
    rs = np.random.RandomState(seed + 5)
    ids = np.sort(rs.choice(H * W, R, replace=False))
    batch = ray_batch_for(scene, ids)
    pc = O.random_ibrnet_params(S_c, seed + 1, sigma_bias=0.3)
    pf = O.random_ibrnet_params(S_c + n_imp, seed + 2, sigma_bias=0.3)
    for p in (pc, pf):      # non-zero biases everywhere
        g = torch.Generator().manual_seed(seed + 9)
        for k in p:
            if k.endswith('.bias'):
                p[k] = p[k] + 0.05 * torch.randn(p[k].shape, generator=g)
    fm = scene['featmaps']

    # ---------- oracle, stage by stage (coarse level) ----------
    t0 = time.time()
    pts, z = O.coarse_depths(batch['ray_o'], batch['ray_d'], batch['depth_range'], S_c, inv_uniform, det=True)
    fm_c = fm[0].clone().requires_grad_(True)
    imgs_req = batch['src_rgbs'].clone().requires_grad_(True)
    rgb_feat, ray_diff, mask = O.projector_compute(pts, batch['camera'], imgs_req, batch['src_cameras'], fm_c)
    want = {}
    rf_leaf = rgb_feat.detach().clone().requires_grad_(True)
    raw = O.ibrnet_forward(pc, pc['pos_encoding'], rf_leaf, ray_diff, mask, True, want=want)
    print(f'oracle coarse stages {time.time() - t0:.1f}s', flush=True)

    # ---------- CUDA: coarse depths ----------
    z_g = ops.coarse_depths(R, S_c, 2.0 if kind == 'llff' else 2.0, float(batch['depth_range'][0, 1]), inv_uniform, None, dev)
    rec(f'{tag}/coarse_z', z_g, z, exact=True)
    tr = torch.rand(R, S_c, generator=torch.Generator().manual_seed(3))
    _, z_j = O.coarse_depths(batch['ray_o'], batch['ray_d'], batch['depth_range'], S_c, inv_uniform, det=False, t_rand=tr)
    z_jg = ops.coarse_depths(R, S_c, float(batch['depth_range'][0, 0]), float(batch['depth_range'][0, 1]), inv_uniform, tr.to(dev), dev)
    rec(f'{tag}/coarse_z_jitter', z_jg, z_j, exact=True)

    # ---------- CUDA: Projector.compute ----------
    proj = Projector(dev)
    fm_cg = fm[0].to(dev).requires_grad_(True)
    imgs_g = batch['src_rgbs'].to(dev).requires_grad_(True)
    cam_g, scam_g = batch['camera'].to(dev), batch['src_cameras'].to(dev)
    rf_g, rd_g, mk_g = proj.compute(pts.to(dev), cam_g, imgs_g, scam_g, fm_cg)
    rec(f'{tag}/proj.mask', mk_g, mask, exact=True)
    rec(f'{tag}/proj.rgb_feat', rf_g, rgb_feat)
    rec(f'{tag}/proj.ray_diff', rd_g, ray_diff)
    cot = torch.randn(rgb_feat.shape, generator=torch.Generator().manual_seed(4))
    (rgb_feat * cot).sum().backward()
    (rf_g * cot.to(dev)).sum().backward()
    rec(f'{tag}/proj.d_featmaps', fm_cg.grad, fm_c.grad)
    rec(f'{tag}/proj.d_imgs', imgs_g.grad, imgs_req.grad)

    # ---------- CUDA: IBRNet view stage / ray stage on the oracle's inputs ----------
    blob = pack_params(pc, device=dev)
    N = R * S_c
    rf_in, rd_in, mk_in = rgb_feat.detach().to(dev).contiguous(), ray_diff.to(dev).contiguous(), mask.to(dev).contiguous()
    ps = torch.zeros(N, 72, device=dev)
    st = _lib.stream_ptr(dev)
    _lib.call('nfb_ibrnet_view_fwd', N, S_c, V, 1, _lib.ptr(rf_in), _lib.ptr(rd_in), _lib.ptr(mk_in), 0, 0, 0, 0,
              None, None, None, None, None, None, None, _lib.ptr(blob), _lib.ptr(ps), None, _lib.precision_code(), st)
    torch.cuda.synchronize()
    ps_c = ps.cpu().view(R, S_c, 72)
    rec(f'{tag}/view.mean2', ps_c[..., 0:32], want['gf_in'][..., 0:32])
    rec(f'{tag}/view.var2', ps_c[..., 32:64], want['gf_in'][..., 32:64])
    rec(f'{tag}/view.wmean', ps_c[..., 64], want['gf_in'][..., 64])
    rec(f'{tag}/view.rgb_out', ps_c[..., 65:68], want['rgb_out'])
    rec(f'{tag}/view.n_valid', ps_c[..., 68], want['n_valid'][..., 0], exact=True)
    # ray stage fed with the ORACLE's per-sample tensor (isolates the ray stage)
    ps_o = torch.zeros(R, S_c, 72)
    ps_o[..., :65] = want['gf_in']
    ps_o[..., 65:68] = want['rgb_out']
    ps_o[..., 68] = want['n_valid'][..., 0]
    ps_og = ps_o.view(N, 72).to(dev).contiguous()
    raw_g = torch.zeros(R, S_c, 4, device=dev)
    pe = pc['pos_encoding'][0].to(dev).contiguous()
    _lib.call('nfb_ibrnet_ray_fwd', R, S_c, _lib.ptr(ps_og), _lib.ptr(blob), _lib.ptr(pe), _lib.ptr(raw_g), None, None, _lib.precision_code(), st)
    torch.cuda.synchronize()
    rec(f'{tag}/ray.sigma(oracle ps)', raw_g[..., 3], raw[..., 3])
    # module forward + backward
    net_c = make_net(pc, S_c)
    rf_gl = rf_in.clone().requires_grad_(True)
    raw_m = net_c(rf_gl, rd_in, mk_in)
    rec(f'{tag}/ibrnet.raw_rgb', raw_m[..., :3], raw[..., :3])
    rec(f'{tag}/ibrnet.raw_sigma', raw_m[..., 3], raw[..., 3])
    cot2 = torch.randn(raw.shape, generator=torch.Generator().manual_seed(6))
    (raw * cot2).sum().backward()
    (raw_m * cot2.to(dev)).sum().backward()
    rec(f'{tag}/ibrnet.d_rgb_feat', rf_gl.grad, rf_leaf.grad)
    # sigma-only and rgb-only cotangents (separates ray-stage and blend gradients)
    for nm, sel in (('sigma', 3), ('rgb', 0)):
        c3 = torch.zeros_like(cot2)
        if sel == 3:
            c3[..., 3] = cot2[..., 3]
        else:
            c3[..., :3] = cot2[..., :3]
        rl = rgb_feat.detach().clone().requires_grad_(True)
        (O.ibrnet_forward(pc, pc['pos_encoding'], rl, ray_diff, mask, True) * c3).sum().backward()
        rg = rf_in.clone().requires_grad_(True)
        (net_c(rg, rd_in, mk_in) * c3.to(dev)).sum().backward()
        rec(f'{tag}/ibrnet.d_rgb_feat[{nm} cot]', rg.grad, rl.grad)

    # ---------- composite ----------
    pixel_mask = mask[..., 0].sum(dim=2) > 1
    raw_l = raw.detach().clone().requires_grad_(True)
    oc = O.composite(raw_l, z, pixel_mask, white_bkgd=False)
    raw_gl = raw.detach().to(dev).requires_grad_(True)
    og = RR.raw2outputs(raw_gl, z.to(dev), pixel_mask.to(dev), white_bkgd=False)
    for k in ('rgb', 'depth', 'weights', 'alpha'):
        rec(f'{tag}/composite.{k}', og[k], oc[k])
    rec(f'{tag}/composite.mask', og['mask'].float(), oc['mask'].float(), exact=True)
    gens = torch.Generator().manual_seed(8)
    cr, cd, cw, ca = (torch.randn(oc[k].shape, generator=gens) for k in ('rgb', 'depth', 'weights', 'alpha'))
    (oc['rgb'] * cr).sum().add((oc['depth'] * cd).sum()).add((oc['weights'] * cw).sum()).add((oc['alpha'] * ca).sum()).backward()
    (og['rgb'] * cr.to(dev)).sum().add((og['depth'] * cd.to(dev)).sum()).add((og['weights'] * cw.to(dev)).sum()).add(
        (og['alpha'] * ca.to(dev)).sum()).backward()
    rec(f'{tag}/composite.d_raw', raw_gl.grad, raw_l.grad)

    # ---------- fine depths ----------
    if n_imp > 0:
        w_c = oc['weights'].detach()
        zf = O.fine_depths(z, w_c, n_imp, inv_uniform, det=True)
        zf_g = RR._fine_z(z.to(dev), w_c.to(dev), n_imp, inv_uniform, True)
        rec(f'{tag}/fine_z(det)', zf_g, zf, exact=True)
        uu = torch.rand(R, n_imp, generator=torch.Generator().manual_seed(10))
        zf2 = O.fine_depths(z, w_c, n_imp, inv_uniform, det=False, u=uu)
        zf2_g = ops.fine_depths(z.to(dev), w_c.to(dev), uu.to(dev), inv_uniform)
        rec(f'{tag}/fine_z(rand u)', zf2_g, zf2, exact=True)

    # ---------- end to end render_rays (fused and composed) + gradients ----------
    t0 = time.time()
    fm_o = (fm[0].clone().requires_grad_(True), fm[1].clone().requires_grad_(True))
    ro = O.render_rays(batch, pc, pf, fm_o, S_c, inv_uniform, n_imp, det=True)
    lo = O.attack_loss(ro, batch['rgb'])
    lo.backward()
    print(f'oracle render_rays+backward {time.time() - t0:.1f}s', flush=True)
    gb = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in batch.items()}
    model = types.SimpleNamespace(net_coarse=make_net(pc, S_c), net_fine=make_net(pf, S_c + n_imp) if n_imp else None)
    for mode in ('fused', 'composed'):
        os.environ['NFB_FUSED'] = '1' if mode == 'fused' else '0'
        fm_g = (fm[0].to(dev).requires_grad_(True), fm[1].to(dev).requires_grad_(True))
        rg = RR.render_rays(gb, model, fm_g, proj, S_c, inv_uniform, n_imp, det=True)
        for lvl in ('coarse', 'fine'):
            if rg['outputs_' + lvl] is None:
                continue
            for k in ('rgb', 'depth', 'weights', 'alpha', 'z_vals'):
                rec(f'{tag}/{mode}.{lvl}.{k}', rg['outputs_' + lvl][k], ro['outputs_' + lvl][k])
            rec(f'{tag}/{mode}.{lvl}.mask', rg['outputs_' + lvl]['mask'].float(), ro['outputs_' + lvl]['mask'].float(), exact=True)
        lg = O.attack_loss({k: ({kk: vv for kk, vv in v.items()} if v is not None else None) for k, v in rg.items()}, gb['rgb'])
        lg.backward()
        rec(f'{tag}/{mode}.loss', lg, lo)
        rec(f'{tag}/{mode}.d_feat_coarse', fm_g[0].grad, fm_o[0].grad)
        if n_imp:
            rec(f'{tag}/{mode}.d_feat_fine', fm_g[1].grad, fm_o[1].grad)
    os.environ['NFB_FUSED'] = '1'


if __name__ == '__main__':
    torch.manual_seed(0)
    print(torch.cuda.get_device_name(0), flush=True)
    quick = '--quick' in sys.argv
    run(V=4, R=256, S_c=64, n_imp=64)
    if not quick:
        run(V=3, R=96, S_c=16, n_imp=16, H=60, W=80, seed=3)
        run(V=10, R=128, S_c=64, n_imp=128, H=200, W=200, kind='synthetic', inv_uniform=False, seed=5)
        run(V=5, R=130, S_c=33, n_imp=31, H=96, W=128, seed=7)
    os.makedirs(os.path.join(REPO, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(REPO, 'gpurun_out', 'diag.json'), 'w') as f:
        json.dump(out, f, indent=1)
    print('DIAG DONE', flush=True)
