"""Two-rank NCCL run of the view-sharded encoder step (attack.delta_gradient_step) with UNEVEN view shards (V = 3 -> 2 + 1:
padded reduce-scatter / all-gathers) against the single-GPU step.  Needs two GPUs (skipped otherwise; the CPU twin over gloo is
tests/test_host_cpu.py::test_delta_gradient_step_with_view_sharded_encoder_gloo)."""
import os
import sys
import types

import numpy as np
import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu


def _setup(dev, V=3, R=256, S=32, NI=32):
    sys.path.insert(0, os.path.join(REPO, 'tests'))
    from test_gpu_parity import _scene, _params, _net
    scene, batch = _scene(V, R, 96, 128, 'llff', seed=5)
    torch.manual_seed(1)
    conv, norm = torch.nn.Conv2d(3, 64, 3, stride=4, padding=1).to(dev), torch.nn.InstanceNorm2d(64).to(dev)

    def enc(x):
        y = norm(conv(x))
        return y[:, :32], y[:, 32:]
    model = types.SimpleNamespace(net_coarse=_net(_params(S, 1), S, dev), net_fine=_net(_params(S + NI, 2), S + NI, dev))
    gb = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in batch.items()}
    delta = ((torch.rand(batch['src_rgbs'].shape, generator=torch.Generator().manual_seed(6)) * 2 - 1) * (8. / 255.)).to(dev)
    return enc, model, gb, delta, S, NI


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    sys.path.insert(0, REPO)
    from nerfool_b200 import attack
    from nerfool_b200.projection import Projector
    enc, model, gb, delta, S, NI = _setup(dev)
    lo, hi = attack.shard_slice(gb['ray_o'].shape[0], rank, world)
    shard = dict(gb)
    for k in ('ray_o', 'ray_d', 'rgb'):
        shard[k] = gb[k][lo:hi].contiguous()
    out = {}
    for shard_enc in (True, False):
        loss, dd = attack.delta_gradient_step(enc, model, Projector(dev), shard, delta, S, NI, inv_uniform=True, det=True,
                                              group=dist.group.WORLD, global_norm=True, shard_encoder=shard_enc)
        out[shard_enc] = (loss.item(), dd.cpu().numpy())
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_view_sharded_step_two_ranks_nccl_uneven_views():
    import torch.multiprocessing as mp
    sys.path.insert(0, REPO)
    from nerfool_b200 import attack
    from nerfool_b200.projection import Projector
    dev = torch.device('cuda:0')
    enc, model, gb, delta, S, NI = _setup(dev)
    loss1, dd1 = attack.delta_gradient_step(enc, model, Projector(dev), gb, delta, S, NI, inv_uniform=True, det=True)
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29731, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    ref = dd1.cpu().numpy()
    for rank, out in res:
        for shard_enc, (loss, dd) in out.items():
            assert abs(loss - loss1.item()) < 1e-6, (rank, shard_enc, loss, loss1.item())
            err = np.linalg.norm(dd - ref) / np.linalg.norm(ref)
            assert err < 1e-4, (rank, shard_enc, err)       # float-atomic scatter order + different shard boundaries
