"""the auxiliary attack losses on the device, against the REFERENCE's own functions
(tests/golden/forward_warp.npz, oracle/make_golden_baseline.py: eval/ibrnet/eval_adv.py:32-48,97-197 and train.py:329-340)."""
import types

import numpy as np
import pytest
import torch

from helpers import load_golden, t, maxabs, relerr, report
from oracle import ibrnet_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
    return torch.device('cuda:0')


@pytest.mark.parametrize('tag', ['pos', 'zero'])
@pytest.mark.parametrize('mode', ['s2t', 't2s', 'full'])
def test_forward_warp_equals_reference_loop(dev, tag, mode):
    """The z-buffer splat on the GPU (order-independent atomicMin form; sequential fallback when a depth is 0) reproduces the
    reference's Python loop exactly: which source pixel wins every destination pixel, its depth and colour, the per-ray
    samples and (tar2src) the re-projected ray indices."""
    from nerfool_b200.aux_losses import forward_warp
    g = load_golden('forward_warp')
    depth = t(g['depth'] if tag == 'pos' else g['depth_zeros']).to(dev)[None]
    kw = {'s2t': dict(src2tar=True), 't2s': dict(src2tar=False), 'full': dict(derive_full_image=True)}[mode]
    out = forward_warp(g['sel'], t(g['rgb']).to(dev), depth, t(g['K0']).to(dev), t(g['E0']).to(dev), t(g['K1']).to(dev), t(g['E1']).to(dev), **kw)
    pre = f'{tag}_{mode}_'
    assert out[0].is_cuda and out[1].is_cuda
    # the reference projects on the CPU, we on the GPU (cuBLAS): projected depths agree to an ulp, so the WINNER of every
    # destination pixel is checked through the colour it carries (copied bit for bit from the source pixel) and the depth to 1e-6
    # relative; a destination may differ only where two candidates are an ulp apart
    new, new_ref = out[0].cpu(), t(g[pre + 'new'])
    mism = int((new != new_ref).any(dim=-1).sum())
    report(f'forward_warp {tag} {mode}: winners differ on {mism} / {new.shape[0] * new.shape[1]} destination pixels; '
           f'max depth deviation {maxabs(out[1].cpu(), g[pre + "new_depth"]):.2e}')
    assert mism <= 2
    assert torch.equal((out[1].cpu() == 0), (t(g[pre + 'new_depth']) == 0))          # same holes
    assert maxabs(out[1].cpu(), g[pre + 'new_depth']) < 2e-5 + (8.0 if mism else 0.0)
    assert (out[2].cpu() != t(g[pre + 'rgb_proj'])).any(dim=1).sum() <= 2
    assert maxabs(out[3].cpu(), g[pre + 'depth_proj']) < 2e-5 + (8.0 if mism else 0.0)
    if mode == 't2s':
        assert (np.asarray(out[4]) != g[pre + 'inds_new']).sum() <= 2


def test_depth_losses_match_reference(dev):
    from nerfool_b200.aux_losses import calc_depth_smooth_loss, calc_depth_var
    g = load_golden('forward_warp')
    dm = t(g['ds_depth']).to(dev)
    assert abs(calc_depth_smooth_loss({'depth': dm}, 8).item() - float(g['ds_l2'])) < 1e-3 * float(g['ds_l2'])
    assert abs(calc_depth_smooth_loss({'depth': dm}, 8, 'l1').item() - float(g['ds_l1'])) < 1e-5 * float(g['ds_l1'])
    ret = {'depth': t(g['dv_depth']).to(dev), 'weights': t(g['dv_weights']).to(dev), 'z_vals': t(g['dv_z']).to(dev)}
    assert abs(calc_depth_var(ret).item() - float(g['dv'])) < 1e-5 * float(g['dv'])


def test_depth_losses_backpropagate_through_fused_render(dev):
    """depth-smooth + depth-variance on top of the RGB loss (eval_adv.py:312-510 terms): their gradients reach the feature maps
    through d depth / d weights of the fused level (nfb_composite_bwd -> ray stage -> view stage).  Against autograd of the
    oracle, truth rule."""
    from test_gpu_parity import _scene, _params, _net, _dbl, _within_truth
    from nerfool_b200.aux_losses import calc_depth_smooth_loss, calc_depth_var
    from nerfool_b200.projection import Projector
    from nerfool_b200.render_ray import render_rays
    from nerfool_b200.attack import rgb_loss
    V, S, NI, ps = 4, 32, 32, 4
    scene, batch = _scene(V, 5 * ps * ps, 96, 128, 'llff', seed=3)
    pc, pf = _params(S, 1), _params(S + NI, 2)

    def total(out, gt, f64):
        loss = O.attack_loss(out, gt) if f64 is not None else rgb_loss(out, gt)
        for lvl in ('outputs_coarse', 'outputs_fine'):
            loss = loss + 0.05 * calc_depth_smooth_loss(out[lvl], ps) + 0.1 * calc_depth_var(out[lvl])
        return loss
    fm32 = tuple(f.clone().requires_grad_(True) for f in scene['featmaps'])
    r32 = O.render_rays(batch, pc, pf, fm32, S, True, NI, det=True)
    total(r32, batch['rgb'], True).backward()
    fm64 = tuple(f.double().requires_grad_(True) for f in scene['featmaps'])
    r64 = O.render_rays(_dbl(batch), _dbl(pc), _dbl(pf), fm64, S, True, NI, det=True, fine_z=r32['outputs_fine']['z_vals'].detach().double())
    total(r64, batch['rgb'].double(), True).backward()
    from nerfool_b200 import render_ray as RR
    gb = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in batch.items()}
    model = types.SimpleNamespace(net_coarse=_net(pc, S, dev), net_fine=_net(pf, S + NI, dev))
    fm = tuple(f.to(dev).requires_grad_(True) for f in scene['featmaps'])
    saved = RR._fine_z
    RR._fine_z = lambda z, w, n, iu, det: r32['outputs_fine']['z_vals'].detach().to(dev)
    try:
        out = render_rays(gb, model, fm, Projector(dev), S, inv_uniform=True, N_importance=NI, det=True)
    finally:
        RR._fine_z = saved
    loss = total(out, gb['rgb'], None)
    loss.backward()
    assert abs(loss.item() - total(r32, batch['rgb'], True).item()) < 1e-3 * abs(loss.item())
    for j, lvl in enumerate(('coarse', 'fine')):
        _within_truth(fm[j].grad, fm32[j].grad, fm64[j].grad, 1e-3, f'rgb + depth-smooth + depth-var loss: d featmaps[{lvl}] (relative)', err=relerr)
