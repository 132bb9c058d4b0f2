"""Profiling probe: a few GNT attack steps (gnt.render_rays -> MSE -> d featmaps) at N_rand rays, for
`ncu --metrics gpu__time_duration.sum` launch lists of nfb_gnt_fwd / nfb_gnt_bwd.  usage: python tests/probes/gnt_attack_probe.py [N_rand] [iters]"""
import os
import sys
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from nerfool_b200.gnt import GNT, Projector, render_rays          # noqa: E402
from nerfool_b200.synthetic import make_scene, rays_for_view      # noqa: E402

nr = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device('cuda:0')
H, W, V, S, depth = 378, 504, 8, 64, 4
scene = make_scene(H, W, V, seed=0, kind='llff')
ray_o, ray_d = rays_for_view(scene['camera'][0], H, W)
idx = torch.arange(0, ray_o.shape[0], ray_o.shape[0] // nr)[:nr]
torch.manual_seed(0)
net = GNT(types.SimpleNamespace(netwidth=64, trans_depth=depth), 32, 63, 63, ret_alpha=True).to(dev).eval()
model = types.SimpleNamespace(net_coarse=net, net_fine=None)
b = {'ray_o': ray_o[idx].to(dev), 'ray_d': ray_d[idx].to(dev), 'depth_range': scene['depth_range'].to(dev), 'camera': scene['camera'][0:1].to(dev),
     'src_rgbs': scene['src_rgbs'].to(dev), 'src_cameras': scene['src_cameras'].to(dev)}
fm = [f.to(dev).requires_grad_(True) for f in scene['featmaps']]
tgt = torch.rand(nr, 3, device=dev)
for _ in range(iters):
    out = render_rays(b, model, fm, Projector(dev), S, inv_uniform=True, N_importance=0, det=True, ret_alpha=True, single_net=True)
    loss = ((out['outputs_coarse']['rgb'] - tgt) ** 2).mean()
    g = torch.autograd.grad(loss, fm[0])[0]
torch.cuda.synchronize()
print('loss', float(loss), 'grad max', float(g.abs().max()))
