"""Host-side logic of one PGD hot-path step and its multi-GPU form (rays / target views sharded across
ranks, one allreduce of the feature-map gradient per step).

Mirrors the hot part of ``optimize_adv_perturb`` (/root/reference/eval/ibrnet/eval_adv.py:258-310):
render_rays on a ray batch -> masked MSE on coarse + fine (criterion.py:23-33, utils.py:48-58) ->
backward to the source feature maps (from where cuDNN's encoder backward carries it to ``delta``)."""
from __future__ import annotations

import torch

from .render_ray import render_rays

TINY_NUMBER = 1e-6


def img2mse(x, y, mask=None):
    """utils.py:48-58 (restated: utils.py itself imports matplotlib)."""
    if mask is None:
        return torch.mean((x - y) * (x - y))
    return torch.sum((x - y) * (x - y) * mask.unsqueeze(-1)) / (torch.sum(mask) * x.shape[-1] + TINY_NUMBER)


def rgb_loss(ret, gt_rgb):
    """Criterion(outputs_coarse) + Criterion(outputs_fine) (eval_adv.py:306-310)."""
    loss = img2mse(ret['outputs_coarse']['rgb'], gt_rgb, ret['outputs_coarse']['mask'].float())
    if ret['outputs_fine'] is not None:
        loss = loss + img2mse(ret['outputs_fine']['rgb'], gt_rgb, ret['outputs_fine']['mask'].float())
    return loss


def shard_slice(n_items: int, rank: int, world: int):
    """Contiguous, balanced shard [lo, hi) of n_items for `rank` (first n_items % world ranks get one more)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pgd_hot_step(model, projector, ray_batch, featmaps, N_samples, N_importance, inv_uniform=True, det=True,
                 white_bkgd=False, max_rays=65536, group=None, global_norm=False, want_img_grad=False):
    """One attack step on the rays of `ray_batch` (already this rank's shard).
    Returns (loss, d_feat_coarse, d_feat_fine); with a process group the gradients are summed over ranks
    with ONE allreduce (both levels packed in one buffer) and the loss is averaged.
    Rays are processed in chunks of `max_rays` to bound the size of the per-sample workspaces; the loss of
    each chunk is normalised by the global mask count so the result equals the un-chunked step.
    want_img_grad: also return d loss / d ray_batch['src_rgbs'] (4th value) -- the source colours enter the
    blending directly (mlp_network.py:233,272), the second path from the perturbation to the loss."""
    fm_c = featmaps[0].detach().requires_grad_(True)
    fm_f = featmaps[1].detach().requires_grad_(True)
    imgs = ray_batch['src_rgbs'].detach().requires_grad_(True) if want_img_grad else None
    g_img = [None, None]                 # per level: the two terms have different normalisers
    R = ray_batch['ray_o'].shape[0]
    n_chunks = max(1, (R + max_rays - 1) // max_rays)
    dev = ray_batch['ray_o'].device
    # loss = num_c / (3 den_c + eps) + num_f / (3 den_f + eps) with num/den summed over chunks.  The coarse
    # term reaches only featmaps[0] and the fine term only featmaps[1] (fine depths are detached,
    # render_ray.py:219), so each chunk back-propagates its un-normalised numerators and the two gradients
    # are scaled by their denominators once at the end: identical to the un-chunked step, bounded memory.
    num = torch.zeros(2, device=dev)
    den = torch.zeros(2, device=dev)
    for i in range(n_chunks):
        lo, hi = i * max_rays, min(R, (i + 1) * max_rays)
        chunk = dict(ray_batch)
        for k in ('ray_o', 'ray_d', 'rgb'):
            chunk[k] = ray_batch[k][lo:hi]
        if want_img_grad:
            chunk['src_rgbs'] = imgs
        ret = render_rays(chunk, model, (fm_c, fm_f), projector, N_samples, inv_uniform=inv_uniform,
                          N_importance=N_importance, det=det, white_bkgd=white_bkgd)
        gt = chunk['rgb']
        part = None
        for j, lvl in enumerate(('coarse', 'fine')):
            o = ret['outputs_' + lvl]
            if o is None:
                continue
            m = o['mask'].float()
            t = torch.sum((o['rgb'] - gt) ** 2 * m.unsqueeze(-1))
            num[j] += t.detach()
            den[j] += torch.sum(m)
            if want_img_grad:
                # one backward per level (their graphs are disjoint: the fine depths are detached), so the image
                # gradient of each level can be scaled by its own normaliser at the end
                t.backward()
                g_img[j] = imgs.grad if g_img[j] is None else g_img[j] + imgs.grad
                imgs.grad = None
            else:
                part = t if part is None else part + t
        if part is not None:
            part.backward()
    multi = group is not None and torch.distributed.get_world_size(group) > 1
    if multi and global_norm:
        # rays of ONE target view sharded over ranks: img2mse divides by the mask count of the whole batch
        # (utils.py:58), so the numerators / denominators are summed over ranks first (4 floats)
        nd = torch.cat([num, den])
        torch.distributed.all_reduce(nd, op=torch.distributed.ReduceOp.SUM, group=group)
        num, den = nd[:2], nd[2:]
    scale = 1.0 / (den * 3 + TINY_NUMBER)
    g_c = fm_c.grad * scale[0]
    g_f = fm_f.grad * scale[1] if fm_f.grad is not None else torch.zeros_like(fm_f)
    loss = (num * scale).sum()
    g_i = None
    if want_img_grad:
        g_i = g_img[0] * scale[0]
        if g_img[1] is not None:
            g_i = g_i + g_img[1] * scale[1]
    if multi:
        parts = [g_c.reshape(-1), g_f.reshape(-1), loss.reshape(1)] + ([g_i.reshape(-1)] if want_img_grad else [])
        packed = torch.cat(parts)
        torch.distributed.all_reduce(packed, op=torch.distributed.ReduceOp.SUM, group=group)
        n = g_c.numel()
        g_c = packed[:n].view_as(g_c)
        g_f = packed[n:2 * n].view_as(g_f)
        if want_img_grad:
            g_i = packed[2 * n + 1:].view_as(g_i)
        total = loss if global_norm else packed[2 * n] / torch.distributed.get_world_size(group)
    else:
        total = loss
    return (total, g_c, g_f, g_i) if want_img_grad else (total, g_c, g_f)


def _all_gather_views(x, counts, group):
    """all_gather of per-rank slices along dim 0 with unequal lengths `counts` (padded to the longest slice)."""
    dist = torch.distributed
    m = max(counts)
    pad = x if x.shape[0] == m else torch.cat([x, x.new_zeros((m - x.shape[0],) + tuple(x.shape[1:]))])
    bufs = [torch.empty_like(pad) for _ in counts]
    dist.all_gather(bufs, pad.contiguous(), group=group)
    return torch.cat([b[:c] for b, c in zip(bufs, counts)])


def delta_gradient_step(encoder, model, projector, ray_batch, delta, N_samples, N_importance, inv_uniform=True, det=True,
                        white_bkgd=False, max_rays=65536, group=None, global_norm=False, shard_encoder=True):
    """d loss / d delta of one attack step, end to end (eval_adv.py:258-310; universal form :609-740), multi-GPU:

        adv = src_rgbs + delta  ->  featmaps = encoder(adv)  ->  render_rays(rays of this rank)  ->  masked MSE

    ``ray_batch['src_rgbs']`` holds the CLEAN source images [1,V,H,W,3]; ``delta`` has the same shape and is replicated
    on every rank; ``encoder(x[v,3,H,W]) -> (feat_coarse[v,32,h,w], feat_fine[v,32,h,w])`` is the reference's ResUNet
    (cuDNN; outside this repo).  Rays (or target views) are sharded by the caller as for ``pgd_hot_step``.
    With a process group and ``shard_encoder`` the ENCODER is sharded over the source views (exact,
    the encoder normalises per image): rank g encodes views shard_slice(V, g, world), the feature maps are
    all-gathered, every rank renders its rays against all V views, the feature-map / image gradients are summed over
    ranks (one packed allreduce -- the reduce-scatter of f2 plus the loss, in one collective), each rank back-propagates
    its own views through its encoder shard, and the delta-gradient slices are all-gathered.  Without ``shard_encoder``
    every rank encodes all views (redundant cuDNN work, no all-gathers).  Returns (loss, d_delta [1,V,H,W,3])."""
    dist = torch.distributed
    multi = group is not None and dist.get_world_size(group) > 1
    world = dist.get_world_size(group) if multi else 1
    rank = dist.get_rank(group) if multi else 0
    src = ray_batch['src_rgbs']
    V = src.shape[1]
    sharded = multi and shard_encoder
    lo, hi = shard_slice(V, rank, world) if sharded else (0, V)
    adv = (src + delta).detach()
    with torch.enable_grad():
        x = adv[0, lo:hi].permute(0, 3, 1, 2).contiguous().requires_grad_(True)
        if hi > lo:
            fc, ff = encoder(x)
        else:                                    # more ranks than source views: this rank encodes nothing
            fc = ff = None
    if sharded:
        counts = [shard_slice(V, r, world)[1] - shard_slice(V, r, world)[0] for r in range(world)]
        probe = torch.tensor([0, 0, 0] if fc is None else list(fc.shape[1:]), device=src.device)
        dist.all_reduce(probe, op=dist.ReduceOp.MAX, group=group)
        shape = tuple(int(v) for v in probe)
        empty = src.new_zeros((0,) + shape)
        full_c = _all_gather_views(fc.detach() if fc is not None else empty, counts, group)
        full_f = _all_gather_views(ff.detach() if ff is not None else empty, counts, group)
    else:
        full_c, full_f = fc.detach(), ff.detach()
    batch = dict(ray_batch)
    batch['src_rgbs'] = adv
    loss, g_c, g_f, g_i = pgd_hot_step(model, projector, batch, (full_c, full_f), N_samples, N_importance,
                                       inv_uniform=inv_uniform, det=det, white_bkgd=white_bkgd, max_rays=max_rays,
                                       group=group, global_norm=global_norm, want_img_grad=True)
    d_local = g_i[0, lo:hi].clone()
    if hi > lo:
        torch.autograd.backward([fc, ff], [g_c[lo:hi].to(fc.dtype), g_f[lo:hi].to(ff.dtype)])
        d_local += x.grad.permute(0, 2, 3, 1)
    d_delta = _all_gather_views(d_local, counts, group) if sharded else d_local
    return loss, d_delta.unsqueeze(0)


class GraphedPGDStep:
    """``pgd_hot_step`` for a FIXED ray-batch size, captured once into a CUDA graph and replayed.

    At the reference's ray-batch sizes (``N_rand`` = 512 by default, config.py:55) one attack step is ~14 kernel launches
    plus the autograd bookkeeping around them and the kernels themselves take ~0.3 ms: the step is launch / host bound.
    The graph removes that: per call only the ray batch and the feature maps are copied into the graph's static
    buffers, the captured kernels are replayed, and the results are read from static output tensors (valid until
    the next call).  Single-GPU (a collective inside the capture is not attempted); the library's entry points are
    stream-ordered and allocate nothing, so the capture only involves PyTorch's graph-private allocator.
    """

    def __init__(self, model, projector, ray_batch, featmaps, N_samples, N_importance, inv_uniform=True, det=True,
                 white_bkgd=False, max_rays=65536):
        if not det:
            raise ValueError('GraphedPGDStep needs det=True (the stochastic sampler draws new random numbers per step)')
        self._batch = dict(ray_batch)
        for k in ('ray_o', 'ray_d', 'rgb'):
            self._batch[k] = ray_batch[k].detach().clone()
        self._fm = [f.detach().clone() for f in featmaps]
        args = (model, projector, self._batch, self._fm, N_samples, N_importance)
        kw = dict(inv_uniform=inv_uniform, det=det, white_bkgd=white_bkgd, max_rays=max_rays)
        # warm-up on a side stream (fills the camera-block / depth-range caches, sizes the allocator), then capture
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                pgd_hot_step(*args, **kw)
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._out = pgd_hot_step(*args, **kw)

    def __call__(self, ray_o, ray_d, rgb, featmaps):
        """Same result as pgd_hot_step on a batch with these rays (shapes must equal the captured ones)."""
        self._batch['ray_o'].copy_(ray_o, non_blocking=True)
        self._batch['ray_d'].copy_(ray_d, non_blocking=True)
        self._batch['rgb'].copy_(rgb, non_blocking=True)
        for dst, src in zip(self._fm, featmaps):
            dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self._out
