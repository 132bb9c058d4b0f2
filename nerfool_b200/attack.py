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


def _has_reduce_scatter(group):
    """reduce_scatter exists on NCCL; gloo (the CPU test backend) has none."""
    try:
        return torch.distributed.get_backend(group) == 'nccl'
    except Exception:
        return False


def pgd_hot_step(model, projector, ray_batch, featmaps, N_samples, N_importance, inv_uniform=True, det=True,
                 white_bkgd=False, max_rays=65536, group=None, global_norm=False, want_img_grad=False, scatter_views=None):
    """One attack step on the rays of `ray_batch` (already this rank's shard).
    Returns (loss, d_feat_coarse, d_feat_fine); with a process group ONE allreduce (both levels + the loss packed in
    one buffer) combines the ranks: the SUM when one view's rays are sharded (``global_norm``: every term already carries
    the global normaliser, so the sum is the single-process result), the MEAN of loss and gradients when every rank
    renders its own target view (the universal attack's minibatch of views).
    Rays are processed in chunks of `max_rays` to bound the size of the per-sample workspaces; the loss of
    each chunk is normalised by the global mask count so the result equals the un-chunked step.
    scatter_views: per-rank source-view counts of a view-sharded encoder (delta_gradient_step): every rank then needs the
    summed gradient of ITS views only, so the exchange is ONE reduce-scatter (NCCL; padded to the largest shard, the loss rides
    in each shard) instead of an allreduce; the returned gradients hold valid data in the rank's own view slice only.
    want_img_grad: also return d loss / d ray_batch['src_rgbs'] (4th value).  The gathered source colours enter the
    blending directly (mlp_network.py:233,272); NOTE that the reference's attacks never perturb them (eval_adv.py:292-304
    and train.py:131-143 pass the CLEAN ray batch to render_rays), so this is off unless a caller asks for it."""
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
    if multi and scatter_views is not None and not want_img_grad and _has_reduce_scatter(group):
        # view-sharded encoder: reduce-scatter.  Shard r = [d feat_c | d feat_f] of rank r's views (zero-padded to the largest
        # shard) + the loss; every rank receives the sum over ranks of its own shard.
        dist = torch.distributed
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        per_view = g_c[0].numel()
        m = max(scatter_views)
        width = 2 * m * per_view + 1
        send = g_c.new_zeros(world, width)
        lo = 0
        for r, c in enumerate(scatter_views):
            if c:
                send[r, :c * per_view] = g_c[lo:lo + c].reshape(-1)
                send[r, m * per_view:(m + c) * per_view] = g_f[lo:lo + c].reshape(-1)
            send[r, -1] = loss
            lo += c
        recv = g_c.new_empty(width)
        dist.reduce_scatter_tensor(recv, send.reshape(-1), op=dist.ReduceOp.SUM, group=group)
        if not global_norm:
            recv = recv / world
        mine, lo = scatter_views[rank], sum(scatter_views[:rank])
        g_c, g_f = torch.zeros_like(g_c), torch.zeros_like(g_f)
        if mine:
            g_c[lo:lo + mine] = recv[:mine * per_view].view(mine, *g_c.shape[1:])
            g_f[lo:lo + mine] = recv[m * per_view:(m + mine) * per_view].view(mine, *g_f.shape[1:])
        total = loss if global_norm else recv[-1]
    elif multi:
        world = torch.distributed.get_world_size(group)
        parts = [g_c.reshape(-1), g_f.reshape(-1), loss.reshape(1)] + ([g_i.reshape(-1)] if want_img_grad else [])
        packed = torch.cat(parts)
        torch.distributed.all_reduce(packed, op=torch.distributed.ReduceOp.SUM, group=group)
        # one view sharded over ranks (global_norm): every rank's terms are already normalised by the GLOBAL mask count, so
        # the sum over ranks IS the single-process gradient; its loss was computed from the summed numerators above.
        # One target view per rank (universal attack): loss and gradients are both the MEAN over the views of the step.
        if not global_norm:
            packed = packed / world
        n_c, n_f = g_c.numel(), g_f.numel()
        o_loss, o_img = n_c + n_f, n_c + n_f + 1
        g_c = packed[:n_c].view_as(g_c)
        g_f = packed[n_c:o_loss].view_as(g_f)
        if want_img_grad:
            g_i = packed[o_img:o_img + g_i.numel()].view_as(g_i)
        total = loss if global_norm else packed[o_loss]
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
                        white_bkgd=False, max_rays=65536, group=None, global_norm=False, shard_encoder=True,
                        perturb_colours=False):
    """d loss / d delta of one attack step, end to end (eval_adv.py:258-310; universal form :609-740), multi-GPU:

        featmaps = encoder(src_rgbs + delta)  ->  render_rays(rays of this rank; colours from the CLEAN src_rgbs)  ->  masked MSE

    exactly as the reference: ``optimize_adv_perturb`` encodes ``src_ray_batch['src_rgbs'] + delta`` (eval_adv.py:290) but
    hands the unperturbed ``src_ray_batch`` to ``render_rays`` (:292-304), and so does adversarial training
    (train.py:129-143) -- ``delta`` reaches the loss only through the feature maps.  ``perturb_colours=True`` (NOT the
    reference's attack) additionally gathers the blended colours from ``src_rgbs + delta`` and adds that direct gradient.

    ``ray_batch['src_rgbs']`` holds the CLEAN source images [1,V,H,W,3]; ``delta`` has the same shape and is replicated
    on every rank; ``encoder(x[v,3,H,W]) -> (feat_coarse[v,32,h,w], feat_fine[v,32,h,w])`` is the reference's ResUNet
    (cuDNN; outside this repo).  Rays (or target views) are sharded by the caller as for ``pgd_hot_step``.
    With a process group and ``shard_encoder`` the ENCODER is sharded over the source views (exact,
    the encoder normalises per image): rank g encodes views shard_slice(V, g, world), the feature maps are
    all-gathered, every rank renders its rays against all V views, the feature-map gradients are combined over
    ranks (ONE reduce-scatter on NCCL: each rank receives the summed gradient of its own views plus the loss; an allreduce on
    backends without reduce-scatter), each rank back-propagates
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
        # shape and dtype of the encoder output, agreed over ranks (a rank without views has none of its own); one small
        # allreduce + host read per (source shape, world size), cached: it would otherwise stall the launch queue every step
        key = (tuple(src.shape), str(src.dtype), world, str(src.device))
        if key not in _ENC_SHAPE_CACHE:
            probe = torch.tensor([0, 0, 0, 0] if fc is None else list(fc.shape[1:]) + [_DTYPE_CODE[fc.dtype]], device=src.device)
            dist.all_reduce(probe, op=dist.ReduceOp.MAX, group=group)
            _ENC_SHAPE_CACHE[key] = (tuple(int(v) for v in probe[:3]), _CODE_DTYPE[int(probe[3])])
        shape, dtype = _ENC_SHAPE_CACHE[key]
        # both levels in ONE all-gather (channels concatenated: [v, 2 C, h, w])
        both = torch.cat([fc.detach(), ff.detach()], dim=1) if fc is not None else torch.zeros((0, 2 * shape[0]) + shape[1:], device=src.device, dtype=dtype)
        full = _all_gather_views(both, counts, group)
        full_c, full_f = full[:, :shape[0]], full[:, shape[0]:]
    else:
        full_c, full_f = fc.detach(), ff.detach()
    batch = dict(ray_batch)
    if perturb_colours:
        batch['src_rgbs'] = adv
    out = pgd_hot_step(model, projector, batch, (full_c, full_f), N_samples, N_importance,
                       inv_uniform=inv_uniform, det=det, white_bkgd=white_bkgd, max_rays=max_rays,
                       group=group, global_norm=global_norm, want_img_grad=perturb_colours,
                       scatter_views=counts if sharded else None)
    loss, g_c, g_f = out[:3]
    d_local = out[3][0, lo:hi].clone() if perturb_colours else None
    if hi > lo:
        torch.autograd.backward([fc, ff], [g_c[lo:hi].to(fc.dtype), g_f[lo:hi].to(ff.dtype)])
        g_enc = x.grad.permute(0, 2, 3, 1)
        d_local = g_enc if d_local is None else d_local + g_enc
    elif d_local is None:
        d_local = src.new_zeros((0,) + tuple(src.shape[2:]))
    d_delta = _all_gather_views(d_local.contiguous(), counts, group) if sharded else d_local
    return loss, d_delta.unsqueeze(0)


_DTYPE_CODE = {torch.float32: 1, torch.float16: 2, torch.bfloat16: 3, torch.float64: 4}
_ENC_SHAPE_CACHE = {}
_CODE_DTYPE = {v: k for k, v in _DTYPE_CODE.items()}


class UniversalAttack:
    """The universal (view-generalisable) perturbation loop of eval_adv.py:609-740 with target views sharded over ranks
    (SURVEY.md 8 row f2 / BASELINE configs[2]).

    The reference takes ONE target view per iteration (``for data in train_loader``), computes the loss of N_rand rays of
    that view against ``encoder(src + delta)`` and steps ``delta`` (Adam on ``-grad`` + StepLR, or ``alpha * sign(grad)``),
    then clamps to the eps-ball and to valid colours (:727-728).  Here one iteration takes ONE VIEW PER RANK (a minibatch of
    ``world`` views; world = 1 is the reference's loop): every rank renders the rays of its own view, the encoder is
    sharded over the source views, one allreduce averages the feature-map gradients, and every rank applies the identical
    optimiser step to its replica of ``delta`` (deterministic, so the replicas stay in lock-step without a broadcast).
    """

    def __init__(self, encoder, model, projector, src_ray_batch, N_samples, N_importance, epsilon=8.0, adv_lr=2.0,
                 use_adam=True, adam_lr=1e-3, lr_step_size=100, lr_gamma=0.5, inv_uniform=True, det=True, white_bkgd=False,
                 group=None, max_rays=65536, shard_encoder=True, generator=None):
        self.encoder, self.model, self.projector = encoder, model, projector
        self.src = src_ray_batch
        self.N_samples, self.N_importance = N_samples, N_importance
        self.kw = dict(inv_uniform=inv_uniform, det=det, white_bkgd=white_bkgd, max_rays=max_rays, group=group,
                       global_norm=False, shard_encoder=shard_encoder)
        src_rgbs = src_ray_batch['src_rgbs']
        dev = src_rgbs.device
        self.epsilon = torch.tensor(epsilon / 255., device=dev)
        self.alpha = torch.tensor(adv_lr / 255., device=dev)
        # init_adv_perturb (eval_adv.py:248-254): U(-eps, eps), clamped to valid colours; the same draw on every rank
        delta = torch.empty(src_rgbs.shape, dtype=src_rgbs.dtype).uniform_(-epsilon / 255., epsilon / 255., generator=generator).to(dev)
        self.delta = torch.max(torch.min(delta, 1 - src_rgbs), 0 - src_rgbs).requires_grad_(True)
        self.use_adam = use_adam
        if use_adam:
            self.opt = torch.optim.Adam([self.delta], lr=adam_lr)
            self.scheduler = torch.optim.lr_scheduler.StepLR(self.opt, step_size=lr_step_size, gamma=lr_gamma)
        self.iters = 0

    def step(self, train_ray_batch):
        """One iteration on this rank's rays (``ray_o``, ``ray_d``, ``rgb``, ``camera``, ``depth_range`` of its target view).
        Returns the loss (mean over the ranks' views)."""
        batch = dict(train_ray_batch)
        batch['src_rgbs'], batch['src_cameras'] = self.src['src_rgbs'], self.src['src_cameras']
        loss, grad = delta_gradient_step(self.encoder, self.model, self.projector, batch, self.delta.detach(),
                                         self.N_samples, self.N_importance, **self.kw)
        with torch.no_grad():
            if self.use_adam:                       # eval_adv.py:693-709: Adam ascends through the negated gradient
                self.opt.zero_grad(set_to_none=True)
                self.delta.grad = -grad.to(self.delta.dtype)
                self.opt.step()
                self.scheduler.step()
            else:                                   # :711-716
                self.delta.add_(self.alpha * torch.sign(grad))
            src = self.src['src_rgbs']
            d = torch.max(torch.min(self.delta, self.epsilon), -self.epsilon)          # :727
            self.delta.copy_(torch.max(torch.min(d, 1 - src), 0 - src))                # :728
        self.iters += 1
        return loss


PGDAttack = UniversalAttack      # the view-specific attack (eval_adv.py:810-852) is the same loop on one fixed target view


class GraphedPGDStep:
    """``pgd_hot_step`` for a FIXED ray-batch size, captured once into a CUDA graph and replayed.

    At the reference's ray-batch sizes (``N_rand`` = 512 by default, config.py:55) one attack step is ~14 kernel launches
    plus the autograd bookkeeping around them and the kernels themselves take ~0.3 ms: the step is launch / host bound.
    The graph removes that: per call the inputs are copied into the graph's static buffers, the captured kernels are
    replayed, and the results are read from static output tensors (valid until the next call).

    Everything the step reads lives in buffers OWNED by this object, so nothing is frozen at capture time: rays and target
    colours, feature maps, source images, the camera block (query-camera centre for ray_diff and the source projection
    matrices -- the reference draws a new target view every iteration, eval_adv.py:651-695) and the coarse depths (a
    function of ``depth_range`` only under ``det=True``).  ``__call__`` refreshes whichever of them the caller passes.
    The IBRNet parameters are frozen (attack mode); a changed parameter raises.  Single-GPU (no collective inside the
    capture); the library's entry points are stream-ordered and allocate nothing.
    """

    def __init__(self, model, projector, ray_batch, featmaps, N_samples, N_importance, inv_uniform=True, det=True,
                 white_bkgd=False, max_rays=65536):
        if not det:
            raise ValueError('GraphedPGDStep needs det=True (the stochastic sampler draws new random numbers per step)')
        from . import ops
        dev = ray_batch['ray_o'].device
        self._inv_uniform, self._N_samples = inv_uniform, N_samples
        self._batch = dict(ray_batch)
        for k in ('ray_o', 'ray_d', 'rgb', 'src_rgbs'):
            self._batch[k] = ray_batch[k].detach().clone()
        R = self._batch['ray_o'].shape[0]
        # private copies of what render_rays would otherwise bake into the graph (render_ray.py: nfb_camera_block / nfb_coarse_z)
        self._cam = ops.camera_block(ray_batch['src_cameras'][0], ray_batch['camera'][0], dev).clone()
        near, far = ops.depth_range_pair(ray_batch['depth_range'])
        self._z = ops.coarse_depths(R, N_samples, near, far, inv_uniform, None, dev)
        self._batch['nfb_camera_block'] = self._cam
        self._batch['nfb_coarse_z'] = self._z
        self._fm = [f.detach().clone() for f in featmaps]
        self._nets = [n for n in (getattr(model, 'net_coarse', None), getattr(model, 'net_fine', None)) if n is not None]
        self._param_key = self._params_now()
        args = (model, projector, self._batch, self._fm, N_samples, N_importance)
        kw = dict(inv_uniform=inv_uniform, det=det, white_bkgd=white_bkgd, max_rays=max_rays)
        # warm-up on a side stream (sizes the allocator, builds the parameter blobs), then capture
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                pgd_hot_step(*args, **kw)
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._out = pgd_hot_step(*args, **kw)

    def _params_now(self):
        return tuple((p.data_ptr(), p._version) for n in self._nets for p in n.parameters())

    def __call__(self, ray_o, ray_d, rgb, featmaps, camera=None, depth_range=None, src_rgbs=None, src_cameras=None):
        """Same result as pgd_hot_step on a batch with these rays (shapes must equal the captured ones).  Pass ``camera``
        ([1,34], the new target view) / ``src_cameras`` / ``depth_range`` / ``src_rgbs`` whenever they differ from the
        previous call: they are copied into the graph's static buffers before the replay."""
        from . import ops
        if self._params_now() != self._param_key:
            raise RuntimeError('GraphedPGDStep: the IBRNet parameters changed since the capture (the graph holds their blob); '
                               'build a new GraphedPGDStep')
        self._batch['ray_o'].copy_(ray_o, non_blocking=True)
        self._batch['ray_d'].copy_(ray_d, non_blocking=True)
        self._batch['rgb'].copy_(rgb, non_blocking=True)
        for dst, src in zip(self._fm, featmaps):
            dst.copy_(src, non_blocking=True)
        if camera is not None or src_cameras is not None:
            if camera is not None:
                self._batch['camera'] = camera
            if src_cameras is not None:
                self._batch['src_cameras'] = src_cameras
            self._cam.copy_(ops.camera_block(self._batch['src_cameras'][0], self._batch['camera'][0], self._cam.device))
        if depth_range is not None:
            near, far = ops.depth_range_pair(depth_range)
            self._z.copy_(ops.coarse_depths(self._z.shape[0], self._N_samples, near, far, self._inv_uniform, None, self._z.device))
        if src_rgbs is not None:
            self._batch['src_rgbs'].copy_(src_rgbs, non_blocking=True)
        self.graph.replay()
        return self._out
