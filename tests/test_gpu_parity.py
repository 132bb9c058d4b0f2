"""GPU parity tests: the CUDA path (through the C ABI, via the reference-shaped Python API) against the CPU
oracle on identical seeded inputs, and against the golden vectors produced by the unmodified reference.

Tolerance policy (DESIGN.md "Parity"):
  * bit-exact: coarse / fine depths, in-frustum masks, pixel masks, sample_pdf bin indices;
  * well-conditioned float stages (gather, compositing, ray stage): 1e-5 abs or tighter;
  * IBRNet's anti-alias pooling weights are DIFFERENCES of exponentials (mlp_network.py:236-239) and cancel
    when source views see a point under similar angles, so a 1-ulp change of exp() is amplified up to
    ~1e3x.  The reference's own fp32 result then sits ~1e-3 from an fp64 evaluation of the same formulas.
    Stages downstream of those weights are therefore judged against the fp64 oracle ("truth"): our error
    must be <= max(north-star tolerance, 3 x the fp32 oracle's own error);
  * gradients: 1e-3 relative (north star), same truth rule where the weights are involved.
"""
import os
import types

import numpy as np
import pytest
import torch

from helpers import (load_golden, params_from_golden, batch_from_golden, t, maxabs, relerr, scene_from_golden,
                     sampled_grad_relerr, report)
from oracle import ibrnet_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
    return torch.device('cuda:0')


def _net(p, S, dev):
    from nerfool_b200.mlp_network import IBRNet
    net = IBRNet(types.SimpleNamespace(anti_alias_pooling=1), 32, S)
    net.load_state_dict({k: v.clone() for k, v in p.items()})
    return net.to(dev).eval()


def _dbl(d):
    return {k: (v.double() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in d.items()}


def _scene(V, R, H=378, W=504, kind='llff', seed=0):
    from nerfool_b200.synthetic import make_scene, ray_batch_for
    scene = make_scene(H, W, V, seed=seed, kind=kind)
    ids = np.sort(np.random.RandomState(seed + 5).choice(H * W, R, replace=False))
    return scene, ray_batch_for(scene, ids)


def _params(S, seed):
    p = O.random_ibrnet_params(S, seed, sigma_bias=0.3)
    g = torch.Generator().manual_seed(seed + 100)
    for k in p:
        if k.endswith('.bias'):
            p[k] = p[k] + 0.05 * torch.randn(p[k].shape, generator=g)
    return p


TRUTH_STATS = {'floor': 0, 'x3': 0}     # how many truth-rule checks passed under the north-star floor vs needed the 3x branch


def _within_truth(ours, o32, o64, floor, what, err=maxabs):
    """err(ours, truth) <= max(floor, 3 * err(oracle32, truth)); logs the measured errors and which branch was needed."""
    e_ours = err(ours.detach().cpu(), o64)
    e_ref = err(o32, o64)
    e_vs32 = err(ours.detach().cpu(), o32)
    branch = 'floor' if e_ours <= floor else 'x3'
    TRUTH_STATS[branch] += 1
    report(f'{what}: |ours-fp64| {e_ours:.2e}  |fp32oracle-fp64| {e_ref:.2e}  |ours-fp32oracle| {e_vs32:.2e}  floor {floor:g} -> {branch} '
           f'(totals: floor {TRUTH_STATS["floor"]}, x3 {TRUTH_STATS["x3"]})')
    assert e_ours <= max(floor, 3.0 * e_ref), f'{what}: ours {e_ours:.3e} vs fp32-oracle {e_ref:.3e} (floor {floor})'


# ---------------------------------------------------------------------------------------------------
# depths
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('inv_uniform', [True, False])
@pytest.mark.parametrize('S', [2, 64, 192])
def test_coarse_depths_bit_exact(dev, inv_uniform, S):
    from nerfool_b200.render_ray import sample_along_camera_ray
    from nerfool_b200 import ops
    R = 37
    o, d = torch.randn(R, 3), torch.randn(R, 3)
    dr = torch.tensor([[2.0, 12.0]])
    pts_o, z_o = O.coarse_depths(o, d, dr, S, inv_uniform=inv_uniform, det=True)
    pts_g, z_g = sample_along_camera_ray(o.to(dev), d.to(dev), dr.to(dev), S, inv_uniform=inv_uniform, det=True)
    assert torch.equal(z_g.cpu(), z_o) and torch.equal(pts_g.cpu(), pts_o)
    tr = torch.rand(R, S, generator=torch.Generator().manual_seed(1))
    _, zj = O.coarse_depths(o, d, dr, S, inv_uniform=inv_uniform, det=False, t_rand=tr)
    zjg = ops.coarse_depths(R, S, 2.0, 12.0, inv_uniform, tr.to(dev), dev)
    assert torch.equal(zjg.cpu(), zj)


def test_empty_batch(dev):
    from nerfool_b200 import ops
    assert ops.coarse_depths(0, 64, 2.0, 6.0, True, None, dev).shape == (0, 64)


# ---------------------------------------------------------------------------------------------------
# Projector.compute
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('V,kind,H,W', [(4, 'llff', 378, 504), (10, 'synthetic', 200, 200), (3, 'llff', 61, 83), (1, 'llff', 48, 64)])
def test_projector_forward_backward(dev, V, kind, H, W):
    from nerfool_b200.projection import Projector
    R, S = 301, 48
    scene, batch = _scene(V, R, H, W, kind, seed=V)
    pts, _ = O.coarse_depths(batch['ray_o'], batch['ray_d'], batch['depth_range'], S, inv_uniform=True, det=True)
    fm = scene['featmaps'][0].clone().requires_grad_(True)
    im = batch['src_rgbs'].clone().requires_grad_(True)
    rf, rd, mk = O.projector_compute(pts, batch['camera'], im, batch['src_cameras'], fm)
    fmg = scene['featmaps'][0].to(dev).requires_grad_(True)
    img = batch['src_rgbs'].to(dev).requires_grad_(True)
    rfg, rdg, mkg = Projector(dev).compute(pts.to(dev), batch['camera'].to(dev), img, batch['src_cameras'].to(dev), fmg)
    assert rfg.shape == rf.shape and rdg.shape == rd.shape and mkg.shape == mk.shape
    assert torch.equal(mkg.cpu(), mk), 'in-frustum mask must be identical'
    assert 0.02 < mk.mean() < 0.9999
    assert maxabs(rfg.cpu(), rf) < 2e-6
    assert maxabs(rdg.cpu(), rd) < 1e-5          # unit(a-b) is itself a cancelling difference
    assert maxabs(rdg.cpu()[..., 3], rd[..., 3]) < 2e-7
    cot = torch.randn(rf.shape, generator=torch.Generator().manual_seed(4))
    (rf * cot).sum().backward()
    (rfg * cot.to(dev)).sum().backward()
    assert relerr(fmg.grad.cpu(), fm.grad) < 1e-5
    assert relerr(img.grad.cpu(), im.grad) < 1e-5


def test_projector_golden(dev):
    from nerfool_b200.projection import Projector
    for name in ('render_llff_v3', 'render_synth_v5'):
        g = load_golden(name)
        rf, rd, mk = Projector(dev).compute(t(g['pts_c']).to(dev), t(g['camera']).to(dev), t(g['src_rgbs']).to(dev),
                                            t(g['src_cameras']).to(dev), t(g['feat_c']).to(dev))
        assert torch.equal(mk.cpu(), t(g['mask_c']))
        assert maxabs(rf.cpu(), g['rgb_feat_c']) < 2e-6
        assert maxabs(rd.cpu()[..., 3], g['ray_diff_c'][..., 3]) < 2e-7


# ---------------------------------------------------------------------------------------------------
# IBRNet.forward
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('V,S,kind,H,W,R', [(4, 64, 'llff', 378, 504, 200), (10, 192, 'synthetic', 200, 200, 40),
                                            (5, 33, 'llff', 96, 128, 131), (8, 128, 'llff', 378, 504, 64),
                                            (2, 256, 'llff', 96, 128, 9)])
def test_ibrnet_forward_backward(dev, V, S, kind, H, W, R):
    scene, batch = _scene(V, R, H, W, kind, seed=V + S)
    pts, _ = O.coarse_depths(batch['ray_o'], batch['ray_d'], batch['depth_range'], S, inv_uniform=True, det=True)
    rf, rd, mk = O.projector_compute(pts, batch['camera'], batch['src_rgbs'], batch['src_cameras'], scene['featmaps'][0])
    p = _params(S, 3)
    rl = rf.clone().requires_grad_(True)
    raw32 = O.ibrnet_forward(p, p['pos_encoding'], rl, rd, mk)
    r64 = rf.double().requires_grad_(True)
    raw64 = O.ibrnet_forward(_dbl(p), p['pos_encoding'].double(), r64, rd.double(), mk.double())
    net = _net(p, S, dev)
    rg = rf.to(dev).requires_grad_(True)
    raw = net(rg, rd.to(dev), mk.to(dev))
    assert raw.shape == (R, S, 4)
    _within_truth(raw[..., :3], raw32[..., :3].detach(), raw64[..., :3].detach(), 1e-5, 'raw rgb')
    _within_truth(raw[..., 3], raw32[..., 3].detach(), raw64[..., 3].detach(), 1e-5, 'raw sigma')
    cot = torch.randn(raw32.shape, generator=torch.Generator().manual_seed(6))
    (raw32 * cot).sum().backward()
    (raw64 * cot.double()).sum().backward()
    (raw * cot.to(dev)).sum().backward()
    e_ours, e_ref = relerr(rg.grad.cpu(), r64.grad), relerr(rl.grad, r64.grad)
    assert e_ours <= max(1e-3, 3 * e_ref), (e_ours, e_ref)


def _grad_params(p, dtype):
    """leaf copies of the float parameters of an oracle parameter dict (pos_encoding stays a constant)."""
    q = {}
    for k, v in p.items():
        if k == 'pos_encoding' or not v.is_floating_point():
            q[k] = v.to(dtype) if v.is_floating_point() else v
        else:
            q[k] = v.detach().to(dtype).clone().requires_grad_(True)
    return q


def _check_param_grads(net, p32, p64, what, floor=1e-2):
    """every parameter gradient of our module within max(floor, 3 x the fp32 oracle's own distance) of the fp64 truth (10 x for the scalar `s`)
    (relative L2 per tensor; tensors whose true gradient is ~0 are compared absolutely).  The floor reflects the
    arithmetic of the view-stage weight gradients: tensor-core GEMMs over the row index with bf16-rounded operands and
    fp32 accumulation (nfb_wgrad_tc.cuh) -- the usual training precision, ~2^-9 per operand before averaging."""
    worst = {}
    gnorm = sum(p64[n].grad.norm().item() ** 2 for n, _ in net.named_parameters() if p64[n].grad is not None) ** 0.5
    for name, prm in net.named_parameters():
        assert prm.grad is not None, f'{what}: no gradient for {name}'
        g, g32, g64 = prm.grad.detach().cpu().double(), p32[name].grad.double(), p64[name].grad
        scale = g64.norm().item()
        if scale < 1e-9:
            # e.g. rgb_fc.4.bias: the blending softmax is shift invariant, the true gradient is exactly 0 and what is
            # left is the fp32 rounding of a sum over all rows
            assert (g - g64).norm().item() < max(1e-4, 3 * (g32 - g64).norm().item()), (what, name, 'zero gradient expected')
            continue
        e_ours = ((g - g64).norm() / scale).item()
        e_ref = ((g32 - g64).norm() / scale).item()
        worst[name] = (e_ours, e_ref)
        # the scalar `s` is one cancelling sum over all rows (fp32 atomics in a run-dependent order): relative to ITSELF its error is
        # a heavy-tailed draw for us and for the fp32 oracle alike, relative to the gradient of the whole net it is ~1e-6 (see
        # tests/test_reference_callers.py::test_reference_train_loop_unmodified); tensor-valued gradients keep the 3x rule
        if name == 's':
            ok = e_ours <= max(floor, 10 * e_ref) or (g - g64).norm().item() < 1e-4 * gnorm
        else:
            ok = e_ours <= max(floor, 3 * e_ref)
        assert ok, f'{what}: d {name}: ours {e_ours:.3e}, fp32 oracle {e_ref:.3e}'
    return worst


@pytest.mark.parametrize('V,S,kind,H,W,R,aa', [(4, 64, 'llff', 378, 504, 100, 1), (10, 192, 'synthetic', 200, 200, 20, 1),
                                               (5, 33, 'llff', 96, 128, 67, 1), (3, 40, 'llff', 96, 128, 50, 0),
                                               (2, 256, 'llff', 96, 128, 5, 1)])
def test_ibrnet_parameter_gradients(dev, V, S, kind, H, W, R, aa):
    """Training (train.py:317-327): d loss / d every IBRNet parameter through IBRNet.forward (tensor form) equals autograd
    of the oracle; the data gradient that comes out of the same kernels is checked too."""
    from nerfool_b200.mlp_network import IBRNet
    scene, batch = _scene(V, R, H, W, kind, seed=V + S)
    pts, _ = O.coarse_depths(batch['ray_o'], batch['ray_d'], batch['depth_range'], S, inv_uniform=True, det=True)
    rf, rd, mk = O.projector_compute(pts, batch['camera'], batch['src_rgbs'], batch['src_cameras'], scene['featmaps'][0])
    p = _params(S, 5)
    cot = torch.randn(R, S, 4, generator=torch.Generator().manual_seed(8))
    p32, p64 = _grad_params(p, torch.float32), _grad_params(p, torch.float64)
    r32 = rf.clone().requires_grad_(True)
    (O.ibrnet_forward(p32, p['pos_encoding'], r32, rd, mk, anti_alias_pooling=bool(aa)) * cot).sum().backward()
    r64 = rf.double().requires_grad_(True)
    (O.ibrnet_forward(p64, p['pos_encoding'].double(), r64, rd.double(), mk.double(), anti_alias_pooling=bool(aa))
     * cot.double()).sum().backward()
    net = IBRNet(types.SimpleNamespace(anti_alias_pooling=aa), 32, S)
    net.load_state_dict({k: v.clone() for k, v in p.items() if aa or k != 's'})
    net = net.to(dev).train()
    rg = rf.to(dev).requires_grad_(True)
    raw = net(rg, rd.to(dev), mk.to(dev))
    (raw * cot.to(dev)).sum().backward()
    _check_param_grads(net, p32, p64, f'V={V} S={S}')
    e_ours, e_ref = relerr(rg.grad.cpu(), r64.grad), relerr(r32.grad, r64.grad)
    assert e_ours <= max(1e-3, 3 * e_ref), (e_ours, e_ref)
    # frozen parameters + data gradient only: the module must not produce parameter gradients
    net.zero_grad(set_to_none=True)
    for q in net.parameters():
        q.requires_grad_(False)
    rg2 = rf.to(dev).requires_grad_(True)
    (net(rg2, rd.to(dev), mk.to(dev)) * cot.to(dev)).sum().backward()
    assert all(q.grad is None for q in net.parameters())
    assert relerr(rg2.grad.cpu(), r64.grad) <= max(1e-3, 3 * e_ref)


def test_render_rays_parameter_gradients(dev):
    """One training step of the fused render_rays: masked-MSE loss -> gradients of both networks' parameters AND of
    the feature maps, against the oracle (coarse level vs the fp64 truth, fine level vs the fp32 oracle at the
    oracle's fine depths)."""
    from nerfool_b200.mlp_network import IBRNet
    from nerfool_b200.projection import Projector
    from nerfool_b200 import render_ray as RR
    from nerfool_b200.attack import rgb_loss
    V, R, S, NI = 4, 96, 32, 32
    scene, batch = _scene(V, R, 378, 504, 'llff', seed=3)
    pc, pf = _params(S, 21), _params(S + NI, 22)
    runs = {}
    for dt in (torch.float32, torch.float64):
        qc, qf = _grad_params(pc, dt), _grad_params(pf, dt)
        fm = tuple(f.detach().clone().to(dt).requires_grad_(True) for f in scene['featmaps'])
        b = batch if dt == torch.float32 else _dbl(batch)
        out = O.render_rays(b, qc, qf, fm, S, True, NI, det=True)
        if dt == torch.float64:      # truth for the coarse level only (the fine depths differ between precisions)
            O.masked_mse(out['outputs_coarse']['rgb'], b['rgb'], out['outputs_coarse']['mask'].double()).backward()
        else:
            O.attack_loss(out, b['rgb']).backward()
        runs[dt] = (qc, qf, fm, out)
    gb = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in batch.items()}
    nets = []
    for p, n in ((pc, S), (pf, S + NI)):
        net = IBRNet(types.SimpleNamespace(anti_alias_pooling=1), 32, n)
        net.load_state_dict({k: v.clone() for k, v in p.items()})
        nets.append(net.to(dev).train())
    model = types.SimpleNamespace(net_coarse=nets[0], net_fine=nets[1])
    fm_g = tuple(f.detach().to(dev).requires_grad_(True) for f in scene['featmaps'])
    saved = RR._fine_z
    RR._fine_z = lambda z, w, n, iu, det: runs[torch.float32][3]['outputs_fine']['z_vals'].to(dev)
    try:
        out = RR.render_rays(gb, model, fm_g, Projector(dev), S, inv_uniform=True, N_importance=NI, det=True)
    finally:
        RR._fine_z = saved
    rgb_loss(out, gb['rgb']).backward()
    _check_param_grads(nets[0], runs[torch.float32][0], runs[torch.float64][0], 'coarse net')
    for name, prm in nets[1].named_parameters():          # fine net: fp32 oracle at identical depths
        ref = runs[torch.float32][1][name].grad
        if ref.norm() > 1e-9:
            assert relerr(prm.grad.cpu(), ref) < 8e-3, ('fine net', name, relerr(prm.grad.cpu(), ref))
    e_c = relerr(fm_g[0].grad.cpu(), runs[torch.float64][2][0].grad)
    e_r = relerr(runs[torch.float32][2][0].grad, runs[torch.float64][2][0].grad)
    assert e_c <= max(1e-3, 3 * e_r), (e_c, e_r)
    assert relerr(fm_g[1].grad.cpu(), runs[torch.float32][2][1].grad) < 5e-3
    # an optimiser step changes the parameters and the next forward sees them (blob cache invalidation)
    before = out['outputs_fine']['rgb'].detach().clone()
    opt = torch.optim.Adam([q for n in nets for q in n.parameters()], lr=1e-2)
    opt.step()
    with torch.no_grad():
        out2 = RR.render_rays(gb, model, fm_g, Projector(dev), S, inv_uniform=True, N_importance=NI, det=True)
    assert maxabs(out2['outputs_fine']['rgb'].cpu(), before.cpu()) > 1e-4


def test_ibrnet_golden_and_no_anti_alias(dev):
    g = load_golden('render_llff_v3')
    p = params_from_golden(g, 'nc')
    net = _net(p, int(g['S_c']), dev)
    rg = t(g['rgb_feat_c']).to(dev).requires_grad_(True)
    raw = net(rg, t(g['ray_diff_c']).to(dev), t(g['mask_c']).to(dev))
    assert maxabs(raw.cpu(), g['raw_c']) < 2e-3       # reference fp32 vs ours: both ~1e-3 from fp64 (see header)
    (raw * t(g['cot_raw_c']).to(dev)).sum().backward()
    assert relerr(rg.grad.cpu(), g['d_rgb_feat_c']) < 2e-3
    # mean pooling variant (args.anti_alias_pooling = 0, mlp_network.py:240-241): well conditioned -> tight
    net.anti_alias_pooling = 0
    raw0 = net(t(g['rgb_feat_c']).to(dev), t(g['ray_diff_c']).to(dev), t(g['mask_c']).to(dev))
    ref0 = O.ibrnet_forward(p, p['pos_encoding'], t(g['rgb_feat_c']), t(g['ray_diff_c']), t(g['mask_c']), anti_alias_pooling=False)
    ref0_64 = O.ibrnet_forward(_dbl(p), p['pos_encoding'].double(), t(g['rgb_feat_c']).double(), t(g['ray_diff_c']).double(),
                               t(g['mask_c']).double(), anti_alias_pooling=False)
    # bf16x3 tensor-core layers + ex2.approx activations: a few 1e-5 on O(1) sigma, well inside the 1e-4 north-star bound
    _within_truth(raw0, ref0, ref0_64, 5e-5, 'raw (mean pooling)')


def test_ibrnet_all_views_masked_and_single_valid(dev):
    """Edge semantics: all-masked samples (uniform blending, sigma forced to 0, masked attention rows) and
    exactly-one-valid-view samples (row mask needs > 1, sigma needs >= 1)."""
    V, S, R = 4, 16, 12
    gen = torch.Generator().manual_seed(0)
    rf = torch.randn(R, S, V, 35, generator=gen)
    rd = torch.randn(R, S, V, 4, generator=gen) * 0.3
    rd[..., 3] = 1 - 0.05 * torch.rand(R, S, V, generator=gen)
    mk = (torch.rand(R, S, V, 1, generator=gen) > 0.4).float()
    mk[0] = 0.0
    mk[1, :, 1:] = 0.0
    mk[1, :, 0] = 1.0
    p = _params(S, 5)
    ref = O.ibrnet_forward(p, p['pos_encoding'], rf, rd, mk)
    ref64 = O.ibrnet_forward(_dbl(p), p['pos_encoding'].double(), rf.double(), rd.double(), mk.double())
    out = _net(p, S, dev)(rf.to(dev), rd.to(dev), mk.to(dev))
    assert torch.all(out[0, :, 3] == 0)
    _within_truth(out, ref, ref64, 2e-5, 'raw (edge cases)')


# ---------------------------------------------------------------------------------------------------
# raw2outputs
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('S,white', [(64, False), (128, True), (200, False), (5, True), (2048, False)])
def test_raw2outputs_forward_backward(dev, S, white):
    from nerfool_b200.render_ray import raw2outputs
    R = 77
    gen = torch.Generator().manual_seed(S)
    raw = torch.rand(R, S, 4, generator=gen)
    raw[..., 3] = torch.relu(torch.randn(R, S, generator=gen)) * 2
    raw[3, :, 3] = 0.0
    raw[4, 2, 3] = 60.0                                   # alpha saturates to 1: transmittance hits the 1e-10 floor
    z = torch.sort(torch.rand(R, S, generator=gen) * 10 + 2, dim=-1)[0]
    pm = torch.rand(R, S, generator=gen) > 0.7
    rl = raw.clone().requires_grad_(True)
    oc = O.composite(rl, z, pm, white_bkgd=white)
    rg = raw.to(dev).requires_grad_(True)
    og = raw2outputs(rg, z.to(dev), pm.to(dev), white_bkgd=white)
    assert list(og.keys()) == ['rgb', 'depth', 'weights', 'mask', 'alpha', 'z_vals']
    assert torch.equal(og['mask'].cpu(), oc['mask']) and og['mask'].dtype == torch.bool
    for k in ('rgb', 'depth', 'weights', 'alpha'):
        assert maxabs(og[k].cpu(), oc[k]) < 5e-6, k
    cots = {k: torch.randn(oc[k].shape, generator=gen) for k in ('rgb', 'depth', 'weights', 'alpha')}
    sum((oc[k] * cots[k]).sum() for k in cots).backward()
    sum((og[k] * cots[k].to(dev)).sum() for k in cots).backward()
    assert relerr(rg.grad.cpu(), rl.grad) < 1e-5
    # only rgb used (the attack loss): other cotangents absent
    rg2 = raw.to(dev).requires_grad_(True)
    raw2outputs(rg2, z.to(dev), pm.to(dev), white_bkgd=white)['rgb'].sum().backward()
    rl2 = raw.clone().requires_grad_(True)
    O.composite(rl2, z, pm, white_bkgd=white)['rgb'].sum().backward()
    assert relerr(rg2.grad.cpu(), rl2.grad) < 1e-5


# ---------------------------------------------------------------------------------------------------
# sample_pdf / fine depths
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('tag', ['det', 'rnd'])
def test_sample_pdf_indices_and_samples_exact(dev, tag):
    from nerfool_b200 import ops
    g = load_golden('sample_pdf')
    bins, w, u = t(g['bins']), t(g['weights']), t(g[f'u_{tag}'])
    s_o, a_o = O.sample_pdf(bins, w, u.shape[1], det=(tag == 'det'), u=u, return_inds=True)
    s_g, a_g = ops.sample_pdf_op(bins.to(dev), w.to(dev), u.to(dev), want_inds=True)
    assert a_g.dtype == torch.int64
    assert torch.equal(a_g.cpu(), a_o), 'bin indices must be identical'
    assert torch.equal(s_g.cpu(), s_o), 'samples must be bit-identical to the oracle'
    # against the reference's own indices (golden): identical up to the documented <=1ulp cdf ties
    assert (a_g.cpu() != t(g[f'above_{tag}'])).float().mean() < 1e-4


def test_sample_pdf_api_mutates_weights_like_reference(dev):
    from nerfool_b200.render_ray import sample_pdf
    g = load_golden('sample_pdf')
    w = t(g['weights']).to(dev)
    w0 = w.clone()
    s = sample_pdf(t(g['bins']).to(dev), w, 64, det=True)
    assert torch.equal(w, w0 + 1e-5)
    assert torch.equal(s.cpu(), O.sample_pdf(t(g['bins']), t(g['weights']), 64, det=True))


@pytest.mark.parametrize('inv_uniform', [True, False])
@pytest.mark.parametrize('S,n_imp', [(64, 64), (64, 128), (16, 7), (3, 5)])
def test_fine_depths_bit_exact(dev, inv_uniform, S, n_imp):
    from nerfool_b200 import ops
    R = 203
    gen = torch.Generator().manual_seed(S + n_imp)
    _, z = O.coarse_depths(torch.zeros(R, 3), torch.ones(R, 3), torch.tensor([[2.0, 12.0]]), S, inv_uniform=inv_uniform, det=True)
    w = torch.rand(R, S, generator=gen) ** 5
    w[:5] = 0
    for det in (True, False):
        u = torch.linspace(0., 1., n_imp) if det else torch.rand(R, n_imp, generator=gen)
        zo = O.fine_depths(z, w, n_imp, inv_uniform=inv_uniform, det=det, u=None if det else u)
        zg = ops.fine_depths(z.to(dev), w.to(dev), u.to(dev), inv_uniform)
        assert torch.equal(zg.cpu(), zo)
        assert bool((zg[:, 1:] >= zg[:, :-1]).all())


# ---------------------------------------------------------------------------------------------------
# render_rays end to end
# ---------------------------------------------------------------------------------------------------
def _render_both(dev, V, R, S, NI, kind, H, W, inv_uniform, white=False, pin_fine_z=False, seed=0):
    """fp32 oracle, fp64 oracle ("truth") and the CUDA path (fused + composed) on the same seeded inputs.
    pin_fine_z: the fp64 truth run and the CUDA runs evaluate the fine level at the fp32 ORACLE's fine depths, so that the
    three differ in arithmetic only; without it every run samples its own fine depths (the end-to-end behaviour)."""
    from nerfool_b200.projection import Projector
    from nerfool_b200 import render_ray as RR
    scene, batch = _scene(V, R, H, W, kind, seed=seed)
    pc, pf = _params(S, seed + 1), _params(S + NI, seed + 2)
    fm_o = tuple(f.clone().requires_grad_(True) for f in scene['featmaps'])
    ro = O.render_rays(batch, pc, pf, fm_o, S, inv_uniform, NI, det=True, white_bkgd=white)
    fm_d = tuple(f.double().requires_grad_(True) for f in scene['featmaps'])
    rd_ = O.render_rays(_dbl(batch), _dbl(pc), _dbl(pf), fm_d, S, inv_uniform, NI, det=True, white_bkgd=white,
                        fine_z=ro['outputs_fine']['z_vals'].double() if pin_fine_z else None)
    gb = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in batch.items()}
    model = types.SimpleNamespace(net_coarse=_net(pc, S, dev), net_fine=_net(pf, S + NI, dev))
    res = {}
    saved = RR._fine_z
    for mode in ('fused', 'composed'):
        os.environ['NFB_FUSED'] = '1' if mode == 'fused' else '0'
        if pin_fine_z:   # test hook: evaluate the fine level at the ORACLE's fine depths
            RR._fine_z = lambda z, w, n, iu, det: ro['outputs_fine']['z_vals'].to(dev)
        try:
            fm_g = tuple(f.to(dev).requires_grad_(True) for f in scene['featmaps'])
            rg = RR.render_rays(gb, model, fm_g, Projector(dev), S, inv_uniform=inv_uniform, N_importance=NI, det=True,
                                white_bkgd=white)
        finally:
            RR._fine_z = saved
            os.environ['NFB_FUSED'] = '1'
        res[mode] = (rg, fm_g)
    return batch, gb, ro, fm_o, rd_, fm_d, res


def _check_fine_depths_end_to_end(z_c, w_ours, w_ref, z_ours, z_ref, NI, inv_uniform, what):
    """End-to-end fine depths (render_ray.py:216-238) of the CUDA pipeline against a reference run.

    (1) WIRING, bit-exact: the pipeline's fine depths equal the oracle's ``fine_depths`` applied to the pipeline's OWN
        coarse weights (detach, [1:-1] slice, flip, inverse-CDF, merge sort -- all on the GPU).
    (2) SENSITIVITY, bounded: against the reference run's depths, which come from ITS coarse weights.  An importance sample
        is ``bin_b + (u - cdf_b) / (cdf_a - cdf_b) * (bin_a - bin_b)``: continuous and piecewise linear in the CDF, so a CDF
        error dC moves a sample of a bin of probability mass p by at most ~2 dC / p bin widths, and never out of the bin's
        neighbourhood.  Bin-index flips may only happen where u is within dC of a CDF entry (ties).
    Returns the report dict."""
    assert torch.equal(z_ours.cpu(), O.fine_depths(z_c, w_ours.cpu(), NI, inv_uniform=inv_uniform, det=True)), what + ': fine-depth wiring'
    R = z_c.shape[0]

    def inv_cdf(w):
        ww = w.clone().detach()[:, 1:-1]
        if inv_uniform:
            iz = 1. / z_c
            bins = torch.flip(.5 * (iz[:, 1:] + iz[:, :-1]), dims=[1])
            ww = torch.flip(ww, dims=[1])
        else:
            bins = .5 * (z_c[:, 1:] + z_c[:, :-1])
        cdf = O.cdf_from_weights(ww)
        u = torch.linspace(0., 1., NI).unsqueeze(0).repeat(R, 1)
        smp, above = O.invert_cdf(bins, cdf, u)
        return bins.double(), cdf.double(), u.double(), smp.double(), above
    bins, cdf_o, u, smp_o, ab_o = inv_cdf(w_ours.cpu())
    _, cdf_r, _, smp_r, ab_r = inv_cdf(w_ref)
    dC = (cdf_o - cdf_r).abs().max(dim=1, keepdim=True)[0] + 2 ** -23                 # per ray, + 1 ulp of the CDF
    below = (ab_r - 1).clamp(min=0)
    mass = (torch.gather(cdf_r, 1, ab_r) - torch.gather(cdf_r, 1, below)).clamp_min(1e-12)
    width = (bins[:, 1:] - bins[:, :-1]).abs()
    wpad = torch.cat([width[:, :1], width, width[:, -1:]], dim=1)                      # neighbourhood width of bin i = max over i-1..i+1
    nb = torch.maximum(torch.maximum(wpad[:, :-2], wpad[:, 1:-1]), wpad[:, 2:])
    bin_id = below.clamp(max=width.shape[1] - 1)
    w_here = torch.gather(nb, 1, bin_id)
    bound = w_here * torch.clamp(4.0 * dC / mass, max=2.0) + 4 * 2 ** -23 * smp_r.abs()
    viol = ((smp_o - smp_r).abs() > bound)
    flips = (ab_o != ab_r)
    # a flip is legitimate only at a tie: u within dC of the CDF entry that separates the two bins
    lo_idx = torch.minimum(ab_o, ab_r)
    tie = (u - torch.gather(cdf_r, 1, lo_idx.clamp(max=cdf_r.shape[1] - 1))).abs() <= 2 * dC
    tie |= (u - torch.gather(cdf_r, 1, (lo_idx + 1).clamp(max=cdf_r.shape[1] - 1))).abs() <= 2 * dC
    bad_flips = int((flips & ~tie).sum())
    rep = O.fine_depth_report(z_ours.cpu(), z_ref)
    rep.update(max_dcdf=float(dC.max()), n_flips=int(flips.sum()), n_flips_not_ties=bad_flips, n_bound_violations=int(viol.sum()),
               max_dw=float((w_ours.cpu() - w_ref).abs().max()))
    report(f'{what}: fine z end to end: max|dz| {rep["max_abs"]:.2e} ({rep["max_rel_range"]:.2e} of the ray\'s depth range), '
           f'{rep["n_exact"]}/{rep["n"]} identical, max|d cdf| {rep["max_dcdf"]:.2e} (max|d w_coarse| {rep["max_dw"]:.2e}), '
           f'bin flips {rep["n_flips"]} of which not at ties {bad_flips}, bound violations {rep["n_bound_violations"]}')
    assert bad_flips == 0, (what, rep)
    assert rep['n_bound_violations'] == 0, (what, rep)
    return rep


def _fine_level_at_own_depths(batch_c, pc, pf, fm_cpu, S, NI, inv_u, white, out, what, gold=None, grads=None):
    """End-to-end parity statement for the fine level when the CUDA pipeline sampled its OWN fine depths:
    (a) the depths themselves are inside the CDF-sensitivity bound of the reference's (asserted by the caller with
        _check_fine_depths_end_to_end), and
    (b) AT those depths the rendering is the reference arithmetic: fp32 / fp64 oracle runs pinned to OUR depths, truth rule with
        the north-star floors (1e-4 RGB / depth, 1e-3 relative for `grads` = (our d featmaps[1])).
    The raw difference to the golden values (which mixes (a) and (b)) is reported."""
    z_own = out['outputs_fine']['z_vals'].detach().cpu()
    fm32 = tuple(f.clone().requires_grad_(True) for f in fm_cpu)
    fm64 = tuple(f.double().requires_grad_(True) for f in fm_cpu)
    r32 = O.render_rays(batch_c, pc, pf, fm32, S, inv_u, NI, det=True, white_bkgd=white, fine_z=z_own)
    r64 = O.render_rays(_dbl(batch_c), _dbl(pc), _dbl(pf), fm64, S, inv_u, NI, det=True, white_bkgd=white, fine_z=z_own.double())
    assert torch.equal(out['outputs_fine']['mask'].cpu(), r32['outputs_fine']['mask'])
    for k in ('rgb', 'depth', 'weights'):
        _within_truth(out['outputs_fine'][k], r32['outputs_fine'][k].detach(), r64['outputs_fine'][k].detach(), 1e-4,
                      f'{what} fine {k} at own depths')
        if gold is not None and ('fine_' + k) in gold:
            report(f'{what} fine {k}: raw |ours - REFERENCE golden| {maxabs(out["outputs_fine"][k].detach().cpu(), gold["fine_" + k]):.2e} '
                   '(includes the sampling shift bounded above)')
    if grads is not None:
        O.attack_loss(r32, batch_c['rgb']).backward()
        O.attack_loss(r64, batch_c['rgb'].double()).backward()
        for j, lvl in enumerate(('coarse', 'fine')):
            _within_truth(grads[j], fm32[j].grad, fm64[j].grad, 1e-3, f'{what} d featmaps[{lvl}] at own depths (relative)', err=relerr)
    return r32, r64


RENDER_CASES = [
    (4, 160, 64, 64, 'llff', 378, 504, True, False),
    (10, 48, 64, 128, 'synthetic', 200, 200, True, True),
    (10, 64, 64, 64, 'llff', 378, 504, True, False),
    (3, 90, 16, 16, 'llff', 61, 83, False, False),
    # contiguous row mapping (a sample's rows straddle warps: V = 7 -> 18 samples per tile, V = 17 -> 7): group-level fences around the
    # cross-view exchanges and around the staging rows of the cooperative gather / scatter
    (7, 44, 16, 16, 'llff', 61, 83, True, False),
    (17, 21, 16, 8, 'llff', 61, 83, True, False),
]


@pytest.mark.parametrize('V,R,S,NI,kind,H,W,inv_uniform,white', RENDER_CASES)
def test_render_rays_outputs(dev, V, R, S, NI, kind, H, W, inv_uniform, white):
    """Both levels against the fp64 truth under the SAME rule (north-star floors: 1e-4 RGB / depth / weights), the fine
    level evaluated at identical depths in all three runs."""
    batch, gb, ro, fm_o, rt, fm_t, res = _render_both(dev, V, R, S, NI, kind, H, W, inv_uniform, white, pin_fine_z=True)
    tag = f'V{V} S{S}+{NI} {H}x{W}'
    for mode, (rg, fm_g) in res.items():
        assert set(rg.keys()) == {'outputs_coarse', 'outputs_fine'}
        for lvl in ('coarse', 'fine'):
            o, r, tr = rg['outputs_' + lvl], ro['outputs_' + lvl], rt['outputs_' + lvl]
            assert list(o.keys()) == ['rgb', 'depth', 'weights', 'mask', 'alpha', 'z_vals']
            assert torch.equal(o['mask'].cpu(), r['mask']), (mode, lvl, 'ray mask')
            assert torch.equal(o['z_vals'].cpu(), r['z_vals']), (mode, lvl, 'depths')
            for k in ('rgb', 'depth', 'weights'):
                _within_truth(o[k], r[k].detach(), tr[k].detach(), 1e-4, f'{tag} {mode} {lvl} {k}')
    # fused and composed paths agree tightly with each other (same kernels, same inputs)
    for lvl in ('coarse', 'fine'):
        for k in ('rgb', 'depth', 'weights', 'alpha'):
            assert maxabs(res['fused'][0]['outputs_' + lvl][k].cpu(), res['composed'][0]['outputs_' + lvl][k].cpu()) < 1e-6


@pytest.mark.parametrize('V,R,S,NI,kind,H,W,inv_uniform,white', RENDER_CASES)
def test_render_rays_fine_depths_end_to_end(dev, V, R, S, NI, kind, H, W, inv_uniform, white):
    """NO pinning: the CUDA pipeline samples its own fine depths from its own coarse weights.  Checks the fine z_vals end to
    end (wiring bit-exact, sensitivity inside the CDF-derived bound, flips only at ties) and the rendered fine RGB against the
    truth rule with the truth evaluated at the oracle's depths (the depth shifts are inside the bound just asserted)."""
    batch, gb, ro, fm_o, rt, fm_t, res = _render_both(dev, V, R, S, NI, kind, H, W, inv_uniform, white, pin_fine_z=False)
    tag = f'V{V} S{S}+{NI} {H}x{W}'
    for mode, (rg, fm_g) in res.items():
        rep = _check_fine_depths_end_to_end(ro['outputs_coarse']['z_vals'], rg['outputs_coarse']['weights'].detach(),
                                            ro['outputs_coarse']['weights'].detach(), rg['outputs_fine']['z_vals'],
                                            ro['outputs_fine']['z_vals'], NI, inv_uniform, f'{tag} {mode}')
        assert torch.equal(rg['outputs_fine']['mask'].cpu(), ro['outputs_fine']['mask'])
        e = maxabs(rg['outputs_fine']['rgb'].cpu(), ro['outputs_fine']['rgb'].detach())
        e_t = maxabs(ro['outputs_fine']['rgb'].detach(), rt['outputs_fine']['rgb'].detach())
        report(f'{tag} {mode}: fine rgb, own fine depths: |ours-fp32oracle| {e:.2e}; |fp32oracle-fp64 run with ITS own depths| {e_t:.2e}')
        assert e <= max(1e-4, 3 * e_t), (mode, e, e_t)


@pytest.mark.parametrize('V,R,S,NI,kind,H,W', [(4, 160, 64, 64, 'llff', 378, 504), (10, 48, 64, 128, 'synthetic', 200, 200),
                                               (10, 64, 64, 64, 'llff', 378, 504)])
def test_render_rays_featmap_gradients(dev, V, R, S, NI, kind, H, W):
    """PGD gradient: d loss / d featmaps of BOTH levels within 1e-3 relative of the fp64 truth, or no worse than 3x the fp32
    oracle's own distance from it (same rule for coarse and fine; all runs at the fp32 oracle's fine depths)."""
    from nerfool_b200.attack import rgb_loss
    batch, gb, ro, fm_o, rt, fm_t, res = _render_both(dev, V, R, S, NI, kind, H, W, True, False, pin_fine_z=True)
    O.attack_loss(ro, batch['rgb']).backward()
    O.attack_loss(rt, batch['rgb'].double()).backward()
    tag = f'V{V} S{S}+{NI} {H}x{W}'
    for mode, (rg, fm_g) in res.items():
        loss = rgb_loss(rg, gb['rgb'])
        loss.backward()
        report(f'{tag} {mode}: loss ours {loss.item():.8f} fp32 oracle {O.attack_loss(ro, batch["rgb"]).item():.8f}')
        assert abs(loss.item() - O.attack_loss(ro, batch['rgb']).item()) < 5e-5
        for j, lvl in enumerate(('coarse', 'fine')):
            _within_truth(fm_g[j].grad, fm_o[j].grad, fm_t[j].grad, 1e-3, f'{tag} {mode} d featmaps[{lvl}] (relative)', err=relerr)


BASE_GOLDENS = ['base_llff_v4', 'base_llff_v10', 'base_synth_v10']


@pytest.mark.parametrize('name', BASE_GOLDENS)
def test_render_rays_baseline_shape_goldens(dev, name):
    """The UNMODIFIED reference's outputs on BASELINE-shaped scenes (378x504 V=4 64+64; 378x504 V=10 64+64; 200x200 V=10
    64+128; oracle/make_golden_baseline.py) against the fused CUDA path, end to end (own fine depths): masks and coarse depths
    identical, fine depths inside the CDF bound, RGB / depth to the north-star 1e-4 and the feature-map gradient to 1e-3
    (truth rule against the fp64 oracle where the fp32 reference itself is further than that from the truth)."""
    from nerfool_b200.projection import Projector
    from nerfool_b200.render_ray import render_rays
    from nerfool_b200.attack import rgb_loss
    g = load_golden(name)
    scene, batch, same = scene_from_golden(g)
    assert same, 'make_scene does not reproduce the inputs this fixture was generated from on this host (digest mismatch)'
    S, NI = int(g['S_c']), int(g['N_imp'])
    pc, pf = params_from_golden(g, 'nc'), params_from_golden(g, 'nf')
    gb = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in batch.items()}
    model = types.SimpleNamespace(net_coarse=_net(pc, S, dev), net_fine=_net(pf, S + NI, dev))
    fm = tuple(f.to(dev).requires_grad_(True) for f in scene['featmaps'])
    out = render_rays(gb, model, fm, Projector(dev), S, inv_uniform=True, N_importance=NI, det=True)
    loss = rgb_loss(out, gb['rgb'])
    loss.backward()
    oc, of = out['outputs_coarse'], out['outputs_fine']
    assert torch.equal(oc['mask'].cpu(), t(g['coarse_mask'])) and torch.equal(of['mask'].cpu(), t(g['fine_mask']))
    assert torch.equal(oc['z_vals'].cpu(), t(g['coarse_z_vals']))
    _check_fine_depths_end_to_end(t(g['coarse_z_vals']), oc['weights'].detach(), t(g['coarse_weights']), of['z_vals'],
                                  t(g['fine_z_vals']), NI, True, name)
    r32, r64 = _fine_level_at_own_depths(batch, pc, pf, scene['featmaps'], S, NI, True, False, out, name, gold=g,
                                         grads=(fm[0].grad, fm[1].grad))
    for k in ('rgb', 'depth', 'weights'):
        _within_truth(oc[k], t(g['coarse_' + k]), r64['outputs_coarse'][k].detach(), 1e-4, f'{name} coarse {k} vs REFERENCE golden')
    report(f'{name}: loss ours {loss.item():.8f} reference {float(g["loss"]):.8f}')
    assert abs(loss.item() - float(g['loss'])) < 1e-4
    # the coarse level's gradient does not depend on the fine depths: directly against the reference's sampled texels
    e_c = sampled_grad_relerr(fm[0].grad, g, 'c')
    n_c = fm[0].grad.double().norm().item() / float(g['d_feat_c_norm'])
    report(f'{name}: d featmaps[coarse] vs REFERENCE golden on the sampled texels: relerr {e_c:.2e}, norm ratio {n_c:.6f}; '
           f'd featmaps[fine] (own depths vs reference depths) relerr {sampled_grad_relerr(fm[1].grad, g, "f"):.2e}')
    assert abs(n_c - 1) < 2e-3


def test_render_rays_source_image_gradient(dev):
    """eval_adv.py perturbs src_rgbs: besides the path through the encoder (d featmaps) the colours enter the
    renderer directly (projection.py:119 -> mlp_network.py:233,272).  d loss / d src_rgbs of the fused path (stash
    and recompute backward) against the fp64 oracle, coarse level (the fine level's depths are not pinned here)."""
    from nerfool_b200 import _lib
    from nerfool_b200.projection import Projector
    from nerfool_b200.render_ray import render_rays
    V, R, S = 4, 150, 48
    scene, batch = _scene(V, R, 96, 128, 'llff', seed=31)
    pc = _params(S, 9)
    b32 = dict(batch); b32['src_rgbs'] = batch['src_rgbs'].clone().requires_grad_(True)
    b64 = _dbl(batch); b64['src_rgbs'] = batch['src_rgbs'].double().requires_grad_(True)
    fm = scene['featmaps']
    r32 = O.render_rays(b32, pc, None, fm, S, True, 0, det=True)
    O.masked_mse(r32['outputs_coarse']['rgb'], batch['rgb'], r32['outputs_coarse']['mask'].float()).backward()
    r64 = O.render_rays(b64, _dbl(pc), None, tuple(f.double() for f in fm), S, True, 0, det=True)
    O.masked_mse(r64['outputs_coarse']['rgb'], batch['rgb'].double(), r64['outputs_coarse']['mask'].double()).backward()
    e_ref = relerr(b32['src_rgbs'].grad, b64['src_rgbs'].grad)
    model = types.SimpleNamespace(net_coarse=_net(pc, S, dev), net_fine=None)
    saved = _lib.STASH_MAX_GIB
    try:
        for cap in (48.0, 0.0):                     # activation-stash backward, then forward-recompute backward
            _lib.STASH_MAX_GIB = cap
            gb = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in batch.items()}
            gb['src_rgbs'] = gb['src_rgbs'].clone().requires_grad_(True)
            out = render_rays(gb, model, tuple(f.to(dev) for f in fm), Projector(dev), S, inv_uniform=True, N_importance=0, det=True)
            m = out['outputs_coarse']['mask'].float()
            loss = torch.sum((out['outputs_coarse']['rgb'] - gb['rgb']) ** 2 * m.unsqueeze(-1)) / (torch.sum(m) * 3 + 1e-6)
            loss.backward()
            e = relerr(gb['src_rgbs'].grad.cpu(), b64['src_rgbs'].grad)
            assert e <= max(1e-3, 3 * e_ref), (cap, e, e_ref)
    finally:
        _lib.STASH_MAX_GIB = saved


def test_render_rays_golden_end_to_end(dev):
    """The reference's own outputs (small golden scenes) through the fused CUDA path, end to end (own fine depths)."""
    from nerfool_b200.projection import Projector
    from nerfool_b200.render_ray import render_rays
    from nerfool_b200.attack import rgb_loss
    for name in ('render_llff_v3', 'render_synth_v5'):
        g = load_golden(name)
        S, NI = int(g['S_c']), int(g['N_imp'])
        inv_u, white = bool(g['inv_uniform']), bool(g['white_bkgd'])
        batch_c = batch_from_golden(g)
        batch = {k: v.to(dev) for k, v in batch_c.items()}
        pc, pf = params_from_golden(g, 'nc'), params_from_golden(g, 'nf')
        model = types.SimpleNamespace(net_coarse=_net(pc, S, dev), net_fine=_net(pf, S + NI, dev))
        fm = (t(g['feat_c']).to(dev).requires_grad_(True), t(g['feat_f']).to(dev).requires_grad_(True))
        out = render_rays(batch, model, fm, Projector(dev), S, inv_uniform=inv_u, N_importance=NI, det=True, white_bkgd=white)
        loss = rgb_loss(out, batch['rgb'])
        loss.backward()
        assert torch.equal(out['outputs_coarse']['mask'].cpu(), t(g['coarse_mask']))
        assert torch.equal(out['outputs_fine']['mask'].cpu(), t(g['fine_mask']))
        assert torch.equal(out['outputs_coarse']['z_vals'].cpu(), t(g['coarse_z_vals']))
        _check_fine_depths_end_to_end(t(g['coarse_z_vals']), out['outputs_coarse']['weights'].detach(), t(g['coarse_weights']),
                                      out['outputs_fine']['z_vals'], t(g['fine_z_vals']), NI, inv_u, name)
        r32, r64 = _fine_level_at_own_depths(batch_c, pc, pf, (t(g['feat_c']), t(g['feat_f'])), S, NI, inv_u, white, out, name, gold=g,
                                             grads=(fm[0].grad, fm[1].grad))
        # coarse level: same depths as the reference -> directly against the golden values
        for k in ('rgb', 'depth', 'weights'):
            _within_truth(out['outputs_coarse'][k], t(g['coarse_' + k]), r64['outputs_coarse'][k].detach(), 1e-4,
                          f'{name} coarse {k} vs REFERENCE golden')
        report(f'{name}: loss ours {loss.item():.8f} reference {float(g["loss"]):.8f}')
        assert abs(loss.item() - float(g['loss'])) < 1e-4
        e_c = relerr(fm[0].grad.cpu(), g['d_feat_c'])
        report(f'{name}: d featmaps[coarse] vs REFERENCE golden relerr {e_c:.2e}')


def test_render_rays_stochastic_sampling_runs_and_is_sorted(dev):
    from nerfool_b200.projection import Projector
    from nerfool_b200.render_ray import render_rays
    scene, batch = _scene(4, 64, 96, 128, 'llff', seed=9)
    gb = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in batch.items()}
    model = types.SimpleNamespace(net_coarse=_net(_params(32, 1), 32, dev), net_fine=_net(_params(48, 2), 48, dev))
    fm = tuple(f.to(dev) for f in scene['featmaps'])
    with torch.no_grad():
        out = render_rays(gb, model, fm, Projector(dev), 32, inv_uniform=True, N_importance=16, det=False)
    z = out['outputs_fine']['z_vals']
    assert z.shape == (64, 48) and bool((z[:, 1:] >= z[:, :-1]).all())
    assert bool(torch.isfinite(out['outputs_fine']['rgb']).all())
    assert float(z.min()) >= 2.0 - 1e-4 and float(z.max()) <= 12.0 + 1e-4


def _small_attack_case(dev, V=4, R=192, S=64, NI=64):
    scene, batch = _scene(V, R, 378, 504, 'llff', seed=21)
    gb = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in batch.items()}
    model = types.SimpleNamespace(net_coarse=_net(_params(S, 7), S, dev), net_fine=_net(_params(S + NI, 8), S + NI, dev))
    return scene, gb, model, S, NI


def _attack_grads(dev, scene, gb, model, S, NI):
    from nerfool_b200.projection import Projector
    from nerfool_b200.render_ray import render_rays
    from nerfool_b200.attack import rgb_loss
    fm = tuple(f.to(dev).requires_grad_(True) for f in scene['featmaps'])
    out = render_rays(gb, model, fm, Projector(dev), S, inv_uniform=True, N_importance=NI, det=True)
    rgb_loss(out, gb['rgb']).backward()
    return out, fm


def test_precision_modes(dev):
    """NfbPrecision: bf16x3 (default) is fp32-equivalent; plain bf16 stays within 0.05 dB PSNR of the fp32 render
    (north star) and its PGD gradient points the same way."""
    from nerfool_b200 import _lib
    scene, gb, model, S, NI = _small_attack_case(dev)
    saved = _lib.get_precision()
    res = {}
    try:
        for mode in ('fp32', 'bf16x3', 'bf16'):
            _lib.set_precision(mode)
            res[mode] = _attack_grads(dev, scene, gb, model, S, NI)
    finally:
        _lib.set_precision(saved)
    ref_out, ref_fm = res['fp32']
    gt = gb['rgb']

    def psnr(o):
        return float(-10.0 * torch.log10(torch.mean((o['outputs_coarse']['rgb'] - gt) ** 2)))

    o3, f3 = res['bf16x3']
    assert maxabs(o3['outputs_coarse']['rgb'].cpu(), ref_out['outputs_coarse']['rgb'].cpu()) < 1e-4
    assert maxabs(o3['outputs_coarse']['depth'].cpu(), ref_out['outputs_coarse']['depth'].cpu()) < 1e-4
    assert relerr(f3[0].grad.cpu(), ref_fm[0].grad.cpu()) < 1e-3
    o1, f1 = res['bf16']
    assert torch.equal(o1['outputs_coarse']['mask'], ref_out['outputs_coarse']['mask'])
    assert abs(psnr(o1) - psnr(ref_out)) < 0.05
    g1, g0 = f1[0].grad.flatten().double(), ref_fm[0].grad.flatten().double()
    assert float(torch.dot(g1, g0) / (g1.norm() * g0.norm())) > 0.995


def test_stash_backward_equals_recompute_backward(dev):
    """The two fused tensor-core backward forms (activation stash vs forward recompute) give the same gradient."""
    from nerfool_b200 import _lib
    scene, gb, model, S, NI = _small_attack_case(dev, V=5, R=77)
    saved = _lib.STASH_MAX_GIB
    try:
        _lib.STASH_MAX_GIB = 48.0
        assert _lib.stash_bytes(77 * S, 5) > 0 and _lib.ray_stash_bytes(77, S) > 0
        _, fm_s = _attack_grads(dev, scene, gb, model, S, NI)
        _lib.STASH_MAX_GIB = 0.0
        assert _lib.stash_bytes(77 * S, 5) == 0
        _, fm_r = _attack_grads(dev, scene, gb, model, S, NI)
    finally:
        _lib.STASH_MAX_GIB = saved
    for a, b in zip(fm_s, fm_r):
        assert relerr(a.grad.cpu(), b.grad.cpu()) < 2e-4


def test_render_single_image_device_resident(dev):
    """render_image.py:21-121 mirror: same dict / shapes / CPU tensors as the reference's chunk loop, equal to one
    un-chunked render_rays over the frame, masked pixels of the coarse image painted white."""
    from nerfool_b200.projection import Projector
    from nerfool_b200.render_ray import render_rays
    from nerfool_b200.render_image import render_single_image
    from nerfool_b200.synthetic import make_scene, ray_batch_for
    Hh, Ww, V, S, NI = 24, 40, 4, 32, 16
    scene = make_scene(Hh, Ww, V, seed=13, kind='llff')
    batch = ray_batch_for(scene, np.arange(Hh * Ww))
    gb = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in batch.items()}
    model = types.SimpleNamespace(net_coarse=_net(_params(S, 3), S, dev), net_fine=_net(_params(S + NI, 4), S + NI, dev))
    fm = tuple(f.to(dev) for f in scene['featmaps'])
    sampler = types.SimpleNamespace(H=Hh, W=Ww)
    os.environ['NFB_RENDER_CHUNK'] = '1'          # keep the caller's 157-ray chunks (default: at least 32768 rays per chunk)
    try:
        with torch.no_grad():
            img = render_single_image(sampler, gb, model, Projector(dev), 157, S, inv_uniform=True, N_importance=NI, det=True,
                                      featmaps=fm)
    finally:
        del os.environ['NFB_RENDER_CHUNK']
    with torch.no_grad():
        big = render_single_image(sampler, gb, model, Projector(dev), 157, S, inv_uniform=True, N_importance=NI, det=True,
                                  featmaps=fm)     # internally one chunk: identical output
        full = render_rays(gb, model, fm, Projector(dev), S, inv_uniform=True, N_importance=NI, det=True)
    for lvl in ('outputs_coarse', 'outputs_fine'):
        for k in img[lvl]:
            assert torch.equal(img[lvl][k], big[lvl][k]), (lvl, k)
    for lvl in ('outputs_coarse', 'outputs_fine'):
        assert list(img[lvl].keys()) == ['rgb', 'depth', 'weights', 'mask', 'alpha', 'z_vals']
        assert img[lvl]['rgb'].shape == (Hh, Ww, 3) and img[lvl]['depth'].shape == (Hh, Ww) and not img[lvl]['rgb'].is_cuda
        assert torch.equal(img[lvl]['depth'], full[lvl]['depth'].cpu().reshape(Hh, Ww))
        assert torch.equal(img[lvl]['mask'], full[lvl]['mask'].cpu().reshape(Hh, Ww))
    want = full['outputs_coarse']['rgb'].cpu().reshape(Hh, Ww, 3).clone()
    want[full['outputs_coarse']['mask'].cpu().reshape(Hh, Ww) == 0] = 1.
    assert torch.equal(img['outputs_coarse']['rgb'], want)
    assert torch.equal(img['outputs_fine']['rgb'], full['outputs_fine']['rgb'].cpu().reshape(Hh, Ww, 3))


def test_render_rays_hybrid(dev):
    """render_ray.py:261-390: with identical clean / adversarial maps the hybrid path is the plain (composed) render;
    with different maps colour and density come from the pass the flags name."""
    from nerfool_b200.projection import Projector
    from nerfool_b200.render_ray import render_rays, render_rays_hybrid, raw2outputs
    scene, batch = _scene(4, 96, 96, 128, 'llff', seed=17)
    gb = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in batch.items()}
    S, NI = 32, 16
    model = types.SimpleNamespace(net_coarse=_net(_params(S, 5), S, dev), net_fine=_net(_params(S + NI, 6), S + NI, dev))
    fm = tuple(f.to(dev) for f in scene['featmaps'])
    fm_adv = tuple(f + 0.05 * torch.randn_like(f) for f in fm)
    proj = Projector(dev)
    with torch.no_grad():
        flags = types.SimpleNamespace(use_clean_color=True, use_clean_density=True)
        same = render_rays_hybrid(gb, model, fm, proj, S, inv_uniform=True, N_importance=NI, det=True, args=flags, featmaps_clean=fm)
        os.environ['NFB_FUSED'] = '0'
        try:
            plain = render_rays(gb, model, fm, proj, S, inv_uniform=True, N_importance=NI, det=True)
        finally:
            os.environ['NFB_FUSED'] = '1'
        for lvl in ('outputs_coarse', 'outputs_fine'):
            for k in ('rgb', 'depth', 'weights', 'mask'):
                assert torch.equal(same[lvl][k], plain[lvl][k]), (lvl, k)
        # colour from the clean maps, density from the adversarial ones (coarse level, checked by hand)
        flags = types.SimpleNamespace(use_clean_color=True, use_clean_density=False)
        mix = render_rays_hybrid(gb, model, fm_adv, proj, S, inv_uniform=True, N_importance=0, det=True, args=flags, featmaps_clean=fm)
        z = mix['outputs_coarse']['z_vals']
        pts = z.unsqueeze(2) * gb['ray_d'].unsqueeze(1) + gb['ray_o'].unsqueeze(1)
        ra = proj.compute(pts, gb['camera'], gb['src_rgbs'], gb['src_cameras'], featmaps=fm_adv[0])
        rc = proj.compute(pts, gb['camera'], gb['src_rgbs'], gb['src_cameras'], featmaps=fm[0])
        raw = torch.cat([model.net_coarse(*rc)[..., :3], model.net_coarse(*ra)[..., 3:4]], dim=2)
        want = raw2outputs(raw, z, ra[2][..., 0].sum(dim=2) > 1)
        assert torch.equal(mix['outputs_coarse']['rgb'], want['rgb'])
        assert not torch.equal(mix['outputs_coarse']['rgb'], same['outputs_coarse']['rgb'])


def _small_scene_gpu(g, dev):
    return {'depth_range': t(g['depth_range']).to(dev), 'camera': t(g['camera'])[:1].to(dev), 'src_rgbs': t(g['src_rgbs']).to(dev),
            'src_cameras': t(g['src_cameras']).to(dev)}


def test_render_single_image_reference_golden(dev):
    """SURVEY 8 row f1 against the REFERENCE: ibrnet/render_image.py:21-123 run by oracle/make_golden_baseline.py on a 40x56
    view (chunk 500) vs the device-resident render_single_image (caller chunk 300 and the default >= 32768-ray chunks)."""
    from nerfool_b200.projection import Projector
    from nerfool_b200.render_image import render_single_image
    from nerfool_b200.synthetic import rays_for_view
    g = load_golden('render_image')
    Hh, Ww, S, NI = int(g['H']), int(g['W']), int(g['S_c']), int(g['N_imp'])
    gb = _small_scene_gpu(g, dev)
    o, d = rays_for_view(t(g['camera'])[0], Hh, Ww)
    gb.update(ray_o=o.to(dev), ray_d=d.to(dev), rgb=None, src_depths=None, depth=None, depth_full=None)
    pc, pf = params_from_golden(g, 'nc'), params_from_golden(g, 'nf')
    model = types.SimpleNamespace(net_coarse=_net(pc, S, dev), net_fine=_net(pf, S + NI, dev))
    fm = (t(g['feat_c']).to(dev), t(g['feat_f']).to(dev))
    sampler = types.SimpleNamespace(H=Hh, W=Ww)
    bc = {k: (v.cpu() if torch.is_tensor(v) else v) for k, v in gb.items()}
    for env in ('1', None):
        if env:
            os.environ['NFB_RENDER_CHUNK'] = env
        try:
            with torch.no_grad():
                img = render_single_image(sampler, gb, model, Projector(dev), 300, S, inv_uniform=True, N_importance=NI, det=True,
                                          featmaps=fm)
        finally:
            os.environ.pop('NFB_RENDER_CHUNK', None)
        for lvl in ('coarse', 'fine'):
            o_ = img['outputs_' + lvl]
            assert list(o_.keys()) == ['rgb', 'depth', 'weights', 'mask', 'alpha', 'z_vals']
            assert all(not v.is_cuda for v in o_.values())
            for k in o_:
                assert o_[k].shape == t(g[f'{lvl}_{k}']).shape and o_[k].dtype == t(g[f'{lvl}_{k}']).dtype, (lvl, k)
            assert torch.equal(o_['mask'], t(g[lvl + '_mask']))
        assert torch.equal(img['outputs_coarse']['z_vals'], t(g['coarse_z_vals']))
        painted = ~t(g['coarse_mask'])
        assert painted.any() and bool((img['outputs_coarse']['rgb'][painted] == 1).all())
        _check_fine_depths_end_to_end(t(g['coarse_z_vals']).reshape(Hh * Ww, -1), img['outputs_coarse']['weights'].reshape(Hh * Ww, -1),
                                      t(g['coarse_weights']).reshape(Hh * Ww, -1), img['outputs_fine']['z_vals'].reshape(Hh * Ww, -1),
                                      t(g['fine_z_vals']).reshape(Hh * Ww, -1), NI, True, 'render_single_image')
        keep = ~painted
        flat = {'outputs_coarse': None, 'outputs_fine': {k: v.reshape(Hh * Ww, *v.shape[2:]) for k, v in img['outputs_fine'].items()}}
        with torch.no_grad():
            r32, r64 = _fine_level_at_own_depths(bc, pc, pf, tuple(f.cpu() for f in fm), S, NI, True, False, flat, 'render_single_image',
                                                 gold={k: (v.reshape(Hh * Ww, *v.shape[2:]) if k.startswith('fine_') else v) for k, v in g.items()})
        for k in ('rgb', 'depth'):
            sel = keep if k == 'rgb' else torch.ones_like(keep)
            shp = (Hh, Ww, 3) if k == 'rgb' else (Hh, Ww)
            _within_truth(img['outputs_coarse'][k][sel], t(g[f'coarse_{k}'])[sel], r64['outputs_coarse'][k].reshape(shp)[sel], 1e-4,
                          f'render_single_image coarse {k} vs REFERENCE golden')


def test_render_rays_hybrid_reference_golden(dev):
    """SURVEY 8 row f4 against the REFERENCE: ibrnet/render_ray.py:261-390 for (use_clean_color, use_clean_density) in
    {(1,0), (0,1), (1,1)} (oracle/make_golden_baseline.py) vs render_rays_hybrid on the GPU."""
    from nerfool_b200.projection import Projector
    from nerfool_b200.render_ray import render_rays_hybrid
    g = load_golden('hybrid')
    S, NI = int(g['S_c']), int(g['N_imp'])
    gb = _small_scene_gpu(g, dev)
    gb.update(ray_o=t(g['ray_o']).to(dev), ray_d=t(g['ray_d']).to(dev))
    pc, pf = params_from_golden(g, 'nc'), params_from_golden(g, 'nf')
    model = types.SimpleNamespace(net_coarse=_net(pc, S, dev), net_fine=_net(pf, S + NI, dev))
    clean = (t(g['feat_c']).to(dev), t(g['feat_f']).to(dev))
    adv = (t(g['adv_c']).to(dev), t(g['adv_f']).to(dev))
    bc = {k: v.cpu() for k, v in gb.items()}
    for cc, cd in ((1, 0), (0, 1), (1, 1)):
        flags = types.SimpleNamespace(use_clean_color=bool(cc), use_clean_density=bool(cd))
        with torch.no_grad():
            out = render_rays_hybrid(gb, model, adv, Projector(dev), S, inv_uniform=True, N_importance=NI, det=True, args=flags,
                                     featmaps_clean=clean)
            o32 = O.render_rays_hybrid(bc, pc, pf, tuple(f.cpu() for f in adv), tuple(f.cpu() for f in clean), S, bool(cc), bool(cd),
                                       inv_uniform=True, n_importance=NI, det=True)
            o64 = O.render_rays_hybrid(_dbl(bc), _dbl(pc), _dbl(pf), tuple(f.cpu().double() for f in adv),
                                       tuple(f.cpu().double() for f in clean), S, bool(cc), bool(cd), inv_uniform=True, n_importance=NI, det=True)
        pre = f'c{cc}d{cd}_'
        for lvl in ('coarse', 'fine'):
            o_ = out['outputs_' + lvl]
            assert torch.equal(o_['mask'].cpu(), t(g[pre + lvl + '_mask'])), (cc, cd, lvl)
        assert torch.equal(out['outputs_coarse']['z_vals'].cpu(), t(g[pre + 'coarse_z_vals']))
        _check_fine_depths_end_to_end(t(g[pre + 'coarse_z_vals']), out['outputs_coarse']['weights'], t(g[pre + 'coarse_weights']),
                                      out['outputs_fine']['z_vals'], t(g[pre + 'fine_z_vals']), NI, True, f'hybrid c{cc}d{cd}')
        for k in ('rgb', 'depth', 'weights'):
            _within_truth(out['outputs_coarse'][k], t(g[pre + 'coarse_' + k]), o64['outputs_coarse'][k], 1e-4,
                          f'hybrid c{cc}d{cd} coarse {k} vs REFERENCE golden')
        # fine level: own fine depths on both sides (differences within the CDF bound asserted above)
        e = maxabs(out['outputs_fine']['rgb'].cpu(), t(g[pre + 'fine_rgb']))
        e32 = maxabs(o32['outputs_fine']['rgb'], t(g[pre + 'fine_rgb']))
        report(f'hybrid c{cc}d{cd} fine rgb: |ours-REFERENCE| {e:.2e}  |fp32 oracle-REFERENCE| {e32:.2e}')
        assert e <= max(1e-4, 3 * e32)


def test_full_size_properties(dev):
    """BASELINE-size chunk (4096 rays, V=4, 64+64): size-independent properties instead of an oracle run:
    weights in [0,1] and sum <= 1, rgb in the convex hull of source colours, determinism, linearity of the
    backward in the upstream gradient, and shard-additivity of the feature-map gradient."""
    from nerfool_b200.projection import Projector
    from nerfool_b200.render_ray import render_rays
    from nerfool_b200.synthetic import make_scene, ray_batch_for
    scene = make_scene(378, 504, 4, seed=0)
    ids = np.arange(100 * 504, 100 * 504 + 4096)
    batch = ray_batch_for(scene, ids)
    gb = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in batch.items()}
    model = types.SimpleNamespace(net_coarse=_net(_params(64, 1), 64, dev), net_fine=_net(_params(128, 2), 128, dev))
    proj = Projector(dev)

    def run(b, scale=1.0):
        fm = tuple(f.to(dev).requires_grad_(True) for f in scene['featmaps'])
        out = render_rays(b, model, fm, proj, 64, inv_uniform=True, N_importance=64, det=True)
        ((out['outputs_coarse']['rgb'] * scale).sum() + (out['outputs_fine']['rgb'] * scale).sum()).backward()
        return out, fm
    out, fm = run(gb)
    for lvl in ('coarse', 'fine'):
        o = out['outputs_' + lvl]
        assert float(o['weights'].min()) >= 0 and float(o['weights'].sum(-1).max()) <= 1 + 1e-5
        assert float(o['rgb'].min()) >= -1e-5 and float(o['rgb'].max()) <= 1 + 1e-5
        assert bool((o['z_vals'][:, 1:] >= o['z_vals'][:, :-1]).all())
    out2, fm2 = run(gb, scale=2.0)
    assert torch.equal(out2['outputs_fine']['rgb'], out['outputs_fine']['rgb'])          # forward is deterministic
    for a, b in zip(fm, fm2):
        assert relerr(b.grad, 2 * a.grad) < 1e-5                                         # backward is linear
    halves = []
    for lo, hi in ((0, 2048), (2048, 4096)):
        hb = dict(gb)
        for k in ('ray_o', 'ray_d', 'rgb'):
            hb[k] = gb[k][lo:hi]
        halves.append(run(hb)[1])
    for lvl in range(2):
        assert relerr(halves[0][lvl].grad + halves[1][lvl].grad, fm[lvl].grad) < 1e-5   # ray shards add up


def test_graft_smoke(dev):
    import __graft_entry__ as ge
    ge.smoke()


def test_graphed_pgd_step_equals_eager(dev):
    """CUDA-graph replay of the attack step (GraphedPGDStep) == the eager step, on new rays and new feature maps."""
    from nerfool_b200.attack import pgd_hot_step, GraphedPGDStep
    from nerfool_b200.projection import Projector
    V, R, S, NI = 4, 512, 64, 64
    scene, batch = _scene(V, 2 * R, 378, 504, 'llff', seed=9)
    model = types.SimpleNamespace(net_coarse=_net(_params(S, 31), S, dev), net_fine=_net(_params(S + NI, 32), S + NI, dev))
    gb = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in batch.items()}
    fm = [f.to(dev) for f in scene['featmaps']]
    first = dict(gb)
    for k in ('ray_o', 'ray_d', 'rgb'):
        first[k] = gb[k][:R].contiguous()
    proj = Projector(dev)
    step = GraphedPGDStep(model, proj, first, fm, S, NI, inv_uniform=True, det=True)
    # replay on the OTHER half of the rays and perturbed feature maps
    second = dict(gb)
    for k in ('ray_o', 'ray_d', 'rgb'):
        second[k] = gb[k][R:].contiguous()
    fm2 = [f + 0.01 * torch.randn_like(f) for f in fm]
    loss_g, gc_g, gf_g = step(second['ray_o'], second['ray_d'], second['rgb'], fm2)
    loss_e, gc_e, gf_e = pgd_hot_step(model, proj, second, fm2, S, NI, inv_uniform=True, det=True)
    assert abs(loss_g.item() - loss_e.item()) < 1e-6
    assert relerr(gc_g.cpu(), gc_e.cpu()) < 1e-5 and relerr(gf_g.cpu(), gf_e.cpu()) < 1e-5     # float-atomic order only
    # and again on the first half: the static buffers are really re-read
    loss_g1, gc_g1, _ = step(first['ray_o'], first['ray_d'], first['rgb'], fm)
    loss_e1, gc_e1, _ = pgd_hot_step(model, proj, first, fm, S, NI, inv_uniform=True, det=True)
    assert abs(loss_g1.item() - loss_e1.item()) < 1e-6 and relerr(gc_g1.cpu(), gc_e1.cpu()) < 1e-5
    assert abs(loss_g1.item() - loss_e.item()) > 1e-6


def test_delta_gradient_step_end_to_end(dev):
    """attack.delta_gradient_step on the GPU (encoder stub -> fused render_rays on the CLEAN source colours -> masked MSE, as
    eval_adv.py:290-304 does) against plain autograd of the oracle with the same encoder on the CPU."""
    from nerfool_b200.attack import delta_gradient_step
    from nerfool_b200.projection import Projector
    V, R, S, NI = 4, 200, 32, 32
    scene, batch = _scene(V, R, 96, 128, 'llff', seed=13)
    pc, pf = _params(S, 41), _params(S + NI, 42)
    torch.manual_seed(5)
    conv, norm = torch.nn.Conv2d(3, 64, 3, stride=4, padding=1), torch.nn.InstanceNorm2d(64)

    def make_enc(c, n):
        def enc(x):
            y = n(c(x))
            return y[:, :32], y[:, 32:]
        return enc
    delta = (torch.rand(batch['src_rgbs'].shape, generator=torch.Generator().manual_seed(6)) * 2 - 1) * (8. / 255.)
    adv = (batch['src_rgbs'] + delta).requires_grad_(True)
    fc, ff = make_enc(conv, norm)(adv[0].permute(0, 3, 1, 2))
    ro = O.render_rays(batch, pc, pf, (fc, ff), S, True, NI, det=True)
    loss0 = O.attack_loss(ro, batch['rgb'])
    loss0.backward()
    import copy
    from nerfool_b200 import render_ray as RR
    model = types.SimpleNamespace(net_coarse=_net(pc, S, dev), net_fine=_net(pf, S + NI, dev))
    gb = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in batch.items()}
    saved = RR._fine_z
    RR._fine_z = lambda z, w, n, iu, det: ro['outputs_fine']['z_vals'].detach().to(dev)
    try:
        loss, dd = delta_gradient_step(make_enc(copy.deepcopy(conv).to(dev), copy.deepcopy(norm).to(dev)), model, Projector(dev), gb,
                                       delta.to(dev), S, NI, inv_uniform=True, det=True, max_rays=256)
    finally:
        RR._fine_z = saved
    assert dd.shape == delta.shape
    assert abs(loss.item() - loss0.item()) < 5e-5
    report(f'delta_gradient_step: loss ours {loss.item():.8f} oracle {loss0.item():.8f}; d delta relerr {relerr(dd.cpu(), adv.grad):.2e}')
    assert relerr(dd.cpu(), adv.grad) < 2e-3, relerr(dd.cpu(), adv.grad)


def test_source_view_permutation_invariance_full_size(dev):
    """BASELINE-size chunks: every cross-view operation of the path (mean / variance pooling, visibility-weighted pooling,
    the blending softmax) is symmetric in the source views, so re-ordering them must leave the rendering unchanged up to
    fp32 summation order, and must permute the feature-map gradient the same way (V = 10, the universal-attack shape)."""
    from nerfool_b200.projection import Projector
    from nerfool_b200.render_ray import render_rays
    from nerfool_b200.synthetic import make_scene, ray_batch_for
    V = 10
    scene = make_scene(378, 504, V, seed=4)
    ids = np.arange(150 * 504, 150 * 504 + 2048)
    batch = ray_batch_for(scene, ids)
    gb = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in batch.items()}
    model = types.SimpleNamespace(net_coarse=_net(_params(64, 1), 64, dev), net_fine=_net(_params(128, 2), 128, dev))
    perm = torch.tensor(np.random.RandomState(0).permutation(V), device=dev)

    def run(order):
        b = dict(gb)
        b['src_rgbs'] = gb['src_rgbs'][:, order].contiguous()
        b['src_cameras'] = gb['src_cameras'][:, order].contiguous()
        fm = tuple(f.to(dev)[order].contiguous().requires_grad_(True) for f in scene['featmaps'])
        out = render_rays(b, model, fm, Projector(dev), 64, inv_uniform=True, N_importance=64, det=True)
        (out['outputs_coarse']['rgb'].sum() + out['outputs_fine']['rgb'].sum()).backward()
        return out, fm
    ident = torch.arange(V, device=dev)
    o1, f1 = run(ident)
    o2, f2 = run(perm)
    assert torch.equal(o1['outputs_coarse']['mask'], o2['outputs_coarse']['mask'])
    for k in ('rgb', 'depth', 'weights'):
        assert maxabs(o1['outputs_coarse'][k].cpu(), o2['outputs_coarse'][k].cpu()) < 2e-5, k
    assert relerr(f2[0].grad.cpu(), f1[0].grad[perm].cpu()) < 1e-4
    # the fine level re-samples from the coarse weights: identical up to samples that sit on a CDF tie
    assert (o1['outputs_fine']['rgb'] - o2['outputs_fine']['rgb']).abs().median().item() < 1e-5


def test_wrapped_nets_train_through_the_wrapper(dev):
    """model.py:78-110 wraps the nets in DistributedDataParallel / DataParallel.  In training the wrapper's own forward
    must run (DDP arms its gradient hooks there), so render_rays takes the composed path through the wrapper; in eval /
    attack mode it keeps the fused path.  Both give the same parameter gradients as the bare modules."""
    from nerfool_b200.mlp_network import IBRNet
    from nerfool_b200.projection import Projector
    from nerfool_b200 import render_ray as RR
    from nerfool_b200.attack import rgb_loss

    class Wrap(torch.nn.Module):             # stands in for DDP / DataParallel: same `.module` attribute, counts forwards
        def __init__(self, m):
            super().__init__()
            self.module, self.calls = m, 0

        def forward(self, *a):
            self.calls += 1
            return self.module(*a)
    V, R, S, NI = 3, 64, 16, 16
    scene, batch = _scene(V, R, 96, 128, 'llff', seed=17)
    gb = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in batch.items()}
    grads = {}
    for wrapped in (False, True):
        nets = []
        for p, n in ((_params(S, 51), S), (_params(S + NI, 52), S + NI)):
            net = IBRNet(types.SimpleNamespace(anti_alias_pooling=1), 32, n)
            net.load_state_dict({k: v.clone() for k, v in p.items()})
            nets.append(net.to(dev).train())
        mods = [Wrap(n) for n in nets] if wrapped else nets
        model = types.SimpleNamespace(net_coarse=mods[0], net_fine=mods[1])
        fm = tuple(f.detach().to(dev).requires_grad_(True) for f in scene['featmaps'])
        assert RR._fusable(model, Projector(dev)) == (not wrapped)
        out = RR.render_rays(gb, model, fm, Projector(dev), S, inv_uniform=True, N_importance=NI, det=True)
        rgb_loss(out, gb['rgb']).backward()
        if wrapped:
            assert mods[0].calls == 1 and mods[1].calls == 1
            for m in mods:
                m.eval()
            assert RR._fusable(model, Projector(dev))          # attack / eval mode: fused path again
        grads[wrapped] = [prm.grad.detach().clone() for n in nets for prm in n.parameters()]
    for a, b in zip(grads[False], grads[True]):
        if a.norm() > 1e-9:
            assert relerr(b.cpu(), a.cpu()) < 2e-2        # two bf16-operand GEMM forms of the same sum (fused vs tensor-mode tiles)


def test_parameter_gradients_with_fully_masked_rays(dev):
    """Rays that see no source view at all (every sample masked: the blending softmax is uniform, sigma is forced to 0)
    contribute exactly what the oracle says to the parameter gradients -- no NaN from the gated tensor-core operands."""
    from nerfool_b200.mlp_network import IBRNet
    V, S, R = 4, 32, 48
    scene, batch = _scene(V, R, 96, 128, 'llff', seed=23)
    pts, _ = O.coarse_depths(batch['ray_o'], batch['ray_d'], batch['depth_range'], S, inv_uniform=True, det=True)
    rf, rd, mk = O.projector_compute(pts, batch['camera'], batch['src_rgbs'], batch['src_cameras'], scene['featmaps'][0])
    mk = mk.clone()
    mk[:16] = 0.            # 16 rays without any valid observation
    mk[16:24, :, 1:] = 0.   # 8 rays with a single valid view
    p = _params(S, 61)
    cot = torch.randn(R, S, 4, generator=torch.Generator().manual_seed(9))
    p32, p64 = _grad_params(p, torch.float32), _grad_params(p, torch.float64)
    (O.ibrnet_forward(p32, p['pos_encoding'], rf, rd, mk) * cot).sum().backward()
    (O.ibrnet_forward(p64, p['pos_encoding'].double(), rf.double(), rd.double(), mk.double()) * cot.double()).sum().backward()
    net = IBRNet(types.SimpleNamespace(anti_alias_pooling=1), 32, S)
    net.load_state_dict({k: v.clone() for k, v in p.items()})
    net = net.to(dev).train()
    (net(rf.to(dev), rd.to(dev), mk.to(dev)) * cot.to(dev)).sum().backward()
    assert all(torch.isfinite(q.grad).all() for q in net.parameters())
    _check_param_grads(net, p32, p64, 'masked rays')
