"""torch.autograd wrappers around the C ABI (one Function per reference operator on the hot path).

PyTorch is plumbing here: it owns device memory and streams and records the autograd graph; every
arithmetic step of the path runs in libnerfool_b200.so.  All functions require CUDA tensors."""
from __future__ import annotations

from collections import OrderedDict

import torch

from . import _lib
from ._lib import call, f32c, ptr, stream_ptr

FEAT_CH = 32
ROW_CH = 35
PS_STRIDE = 72


# --------------------------------------------------------------------------------------------------
# camera block (include/nerfool_b200.h): P_v = K_v inverse(c2w_v) built on the HOST with the reference's
# own torch ops (projection.py:52-56) so the in-frustum masks are bit-identical to the CPU reference.
# --------------------------------------------------------------------------------------------------
_cam_cache: "OrderedDict[tuple, tuple]" = OrderedDict()
_CAM_CACHE_MAX = 32


def _tensor_key(t: torch.Tensor):
    """Identity of the memory a tensor (or a view such as ``src_cameras[0]``, a fresh Python object on every call) looks
    at, plus the version counter views share with their base: equal keys <=> same bytes, as long as the tensor is alive."""
    return (t.untyped_storage().data_ptr(), t.storage_offset(), tuple(t.shape), tuple(t.stride()), t._version, str(t.device))


def camera_block(train_cameras: torch.Tensor, query_camera: torch.Tensor, device, H=None, W=None) -> torch.Tensor:
    """train_cameras [V,34], query_camera [34] (any device) -> device float32 [16*V+4].
    Cached per (storage, view geometry, version); the cache keeps the key tensors alive so addresses cannot be reused.
    H, W: size of the source images the kernels will normalise / bounds-check with.  The reference reads them from the
    camera vector (``h, w = train_cameras[0][:2]``, projection.py:112); the kernels take them from the image tensor, so
    the two must agree -- checked here, once per cached camera tensor."""
    key = (_tensor_key(train_cameras), _tensor_key(query_camera), str(device), H, W)
    hit = _cam_cache.get(key)
    if hit is not None:
        _cam_cache.move_to_end(key)
        return hit[0]
    tc = train_cameras.detach().float().cpu()
    qc = query_camera.detach().float().cpu()
    V = tc.shape[0]
    if H is not None and (float(tc[0, 0]) != float(H) or float(tc[0, 1]) != float(W)):
        raise RuntimeError(f'source camera vector says h, w = {float(tc[0, 0]):g}, {float(tc[0, 1]):g} but the source images are '
                           f'{H} x {W}: the reference normalises projections with the camera\'s h, w (projection.py:112); '
                           'resized source images need matching camera vectors')
    K = tc[:, 2:18].reshape(-1, 4, 4)
    c2w = tc[:, -16:].reshape(-1, 4, 4)
    P = K.bmm(torch.inverse(c2w))                              # [V,4,4], same ops as the reference
    blk = torch.zeros(16 * V + 4, dtype=torch.float32)
    per = blk[:16 * V].view(V, 16)
    per[:, :12] = P[:, :3, :].reshape(V, 12)
    per[:, 12:15] = c2w[:, :3, 3]
    blk[16 * V:16 * V + 3] = qc[-16:].reshape(4, 4)[:3, 3]
    dev_blk = blk.to(device, non_blocking=False)
    _cam_cache[key] = (dev_blk, train_cameras, query_camera)
    while len(_cam_cache) > _CAM_CACHE_MAX:
        _cam_cache.popitem(last=False)
    return dev_blk


_range_cache: "OrderedDict[tuple, tuple]" = OrderedDict()


def depth_range_pair(depth_range: torch.Tensor):
    """(near, far) python floats of a [1,2] tensor; cached per tensor object so a device-resident
    depth_range costs one device->host read, not one per call (render_ray.py:85-87 reads it every call)."""
    key = (id(depth_range), depth_range._version)
    hit = _range_cache.get(key)
    if hit is not None:
        _range_cache.move_to_end(key)
        return hit[0]
    vals = depth_range.detach().float().cpu()
    pair = (float(vals[0, 0]), float(vals[0, 1]))
    _range_cache[key] = (pair, depth_range)
    while len(_range_cache) > _CAM_CACHE_MAX:
        _range_cache.popitem(last=False)
    return pair


def channels_last_feat(featmaps: torch.Tensor) -> torch.Tensor:
    """[V,32,h,w] (any strides) -> contiguous [V,h,w,32] fp32."""
    if featmaps.shape[1] != FEAT_CH:
        raise RuntimeError(f'nerfool_b200 kernels are built for {FEAT_CH} feature channels, got {featmaps.shape[1]}')
    return f32c(featmaps.permute(0, 2, 3, 1))


class _Geom:
    """Geometry arguments shared by the gather / fused kernels."""
    __slots__ = ('N', 'S', 'V', 'H', 'W', 'fh', 'fw', 'xyz', 'ray_o', 'ray_d', 'z', 'cam')

    def ints(self):
        return (self.N, self.S, self.V, self.H, self.W, self.fh, self.fw)

    def pts(self):
        return (ptr(self.xyz), ptr(self.ray_o), ptr(self.ray_d), ptr(self.z))


# --------------------------------------------------------------------------------------------------
# Projector.compute
# --------------------------------------------------------------------------------------------------
class ProjectGather(torch.autograd.Function):
    """rgb_feat, ray_diff, mask = f(xyz | (ray_o, ray_d, z), imgs[V,H,W,3], featmaps[V,32,h,w], cam)."""

    @staticmethod
    def forward(ctx, xyz, imgs, featmaps, cam, H, W):
        _lib.require_cuda(xyz, imgs, featmaps, cam)
        xyz_c = f32c(xyz)
        imgs_c = f32c(imgs)
        feat = channels_last_feat(featmaps)
        R, S = xyz.shape[:2]
        V, fh, fw = feat.shape[0], feat.shape[1], feat.shape[2]
        N = R * S
        dev = xyz.device
        rgb_feat = torch.empty(R, S, V, ROW_CH, device=dev, dtype=torch.float32)
        ray_diff = torch.empty(R, S, V, 4, device=dev, dtype=torch.float32)
        mask = torch.empty(R, S, V, 1, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            call('nfb_project_gather_fwd', N, S, V, H, W, fh, fw, ptr(xyz_c), None, None, None, ptr(cam),
                 ptr(imgs_c), ptr(feat), ptr(rgb_feat), ptr(ray_diff), ptr(mask), stream_ptr(dev))
        ctx.save_for_backward(xyz_c, cam)
        ctx.dims = (N, S, V, H, W, fh, fw)
        ctx.imgs_shape = imgs.shape
        ctx.mark_non_differentiable(ray_diff, mask)
        return rgb_feat, ray_diff, mask

    @staticmethod
    def backward(ctx, d_rgb_feat, _d_rd, _d_mask):
        xyz_c, cam = ctx.saved_tensors
        N, S, V, H, W, fh, fw = ctx.dims
        dev = xyz_c.device
        need_imgs, need_feat = ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        d_feat = torch.zeros(V, fh, fw, FEAT_CH, device=dev, dtype=torch.float32) if need_feat else None
        d_imgs = torch.zeros(ctx.imgs_shape, device=dev, dtype=torch.float32) if need_imgs else None
        if need_feat or need_imgs:
            g = f32c(d_rgb_feat)
            with torch.cuda.device(dev):
                call('nfb_project_gather_bwd', N, S, V, H, W, fh, fw, ptr(xyz_c), None, None, None, ptr(cam),
                     ptr(g), ptr(d_feat), ptr(d_imgs), stream_ptr(dev))
        return (None, d_imgs, d_feat.permute(0, 3, 1, 2) if need_feat else None, None, None, None)


class ProjectGatherCam(torch.autograd.Function):
    """Projector.compute with the source cameras in the graph (gnt/projection.py:84-132 does not detach them; eval/gnt/eval_adv.py
    --perturb_camera optimises source rotations / translations): rgb_feat, ray_diff, mask = f(xyz, imgs, featmaps, train_cameras
    [V,34], query_camera [34]).  Forward = the same kernel as ProjectGather.  Backward: d featmaps / d imgs by the scatter kernel;
    d train_cameras = the grid gradient of both gathers (``nfb_project_grid_bwd``) and the incoming d ray_diff chained through the
    reference's own camera arithmetic (projection.py:42-87: K inverse(c2w), perspective divide, clamps, unit-vector differences)
    replayed with torch autograd -- 34 floats per view, not on the hot path."""

    @staticmethod
    def forward(ctx, xyz, imgs, featmaps, train_cameras, query_camera, H, W):
        _lib.require_cuda(xyz, imgs, featmaps)
        cam = camera_block(train_cameras, query_camera, xyz.device, H, W)
        xyz_c, imgs_c, feat = f32c(xyz), f32c(imgs), channels_last_feat(featmaps)
        R, S = xyz.shape[:2]
        V, fh, fw = feat.shape[0], feat.shape[1], feat.shape[2]
        N, dev = R * S, xyz.device
        rgb_feat = torch.empty(R, S, V, ROW_CH, device=dev, dtype=torch.float32)
        ray_diff = torch.empty(R, S, V, 4, device=dev, dtype=torch.float32)
        mask = torch.empty(R, S, V, 1, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            call('nfb_project_gather_fwd', N, S, V, H, W, fh, fw, ptr(xyz_c), None, None, None, ptr(cam),
                 ptr(imgs_c), ptr(feat), ptr(rgb_feat), ptr(ray_diff), ptr(mask), stream_ptr(dev))
        ctx.save_for_backward(xyz_c, cam, imgs_c, feat, train_cameras.detach(), query_camera.detach())
        ctx.dims = (N, S, V, H, W, fh, fw)
        ctx.imgs_shape = imgs.shape
        ctx.mark_non_differentiable(mask)
        return rgb_feat, ray_diff, mask

    @staticmethod
    def backward(ctx, d_rgb_feat, d_ray_diff, _d_mask):
        xyz_c, cam, imgs_c, feat, cams, qcam = ctx.saved_tensors
        N, S, V, H, W, fh, fw = ctx.dims
        dev = xyz_c.device
        need_imgs, need_feat, need_cam = ctx.needs_input_grad[1], ctx.needs_input_grad[2], ctx.needs_input_grad[3]
        d_feat = torch.zeros(V, fh, fw, FEAT_CH, device=dev, dtype=torch.float32) if need_feat else None
        d_imgs = torch.zeros(ctx.imgs_shape, device=dev, dtype=torch.float32) if need_imgs else None
        g = f32c(d_rgb_feat) if d_rgb_feat is not None else None
        if g is not None and (need_feat or need_imgs):
            with torch.cuda.device(dev):
                call('nfb_project_gather_bwd', N, S, V, H, W, fh, fw, ptr(xyz_c), None, None, None, ptr(cam),
                     ptr(g), ptr(d_feat), ptr(d_imgs), stream_ptr(dev))
        d_cams = None
        if need_cam:
            outs, cots = [], []
            with torch.enable_grad():
                c = cams.to(dev).float().requires_grad_(True)
                flat = xyz_c.reshape(-1, 3)
                c2w = c[:, -16:].reshape(-1, 4, 4)
                if g is not None:
                    d_grid = torch.empty(N, V, 2, device=dev, dtype=torch.float32)
                    with torch.cuda.device(dev):
                        call('nfb_project_grid_bwd', N, S, V, H, W, fh, fw, ptr(xyz_c), None, None, None, ptr(cam),
                             ptr(imgs_c), ptr(feat), ptr(g), ptr(d_grid), stream_ptr(dev))
                    P = c[:, 2:18].reshape(-1, 4, 4).bmm(torch.inverse(c2w))
                    homog = torch.cat([flat, torch.ones_like(flat[:, :1])], dim=-1)
                    proj = P.bmm(homog.t()[None].repeat(V, 1, 1)).permute(0, 2, 1)                     # [V,N,4]
                    pix = torch.clamp(proj[..., :2] / torch.clamp(proj[..., 2:3], min=1e-8), min=-1e6, max=1e6)
                    grid = 2 * pix / torch.tensor([W - 1., H - 1.], device=dev)[None, None, :] - 1.
                    outs.append(grid)
                    cots.append(d_grid.permute(1, 0, 2))
                if d_ray_diff is not None:
                    tgt = qcam.to(dev).float()[-16:].reshape(4, 4)[:3, 3]
                    a = tgt[None, None, :] - flat[None]
                    a = a / (torch.norm(a, dim=-1, keepdim=True) + 1e-6)
                    b = c2w[:, :3, 3].unsqueeze(1) - flat[None]
                    b = b / (torch.norm(b, dim=-1, keepdim=True) + 1e-6)
                    d = a - b
                    rd = torch.cat([d / torch.clamp(torch.norm(d, dim=-1, keepdim=True), min=1e-6), torch.sum(a * b, dim=-1, keepdim=True)], dim=-1)
                    outs.append(rd)                                                                     # [V,N,4]
                    cots.append(f32c(d_ray_diff).reshape(N, V, 4).permute(1, 0, 2))
                if outs:
                    (d_cams,) = torch.autograd.grad(outs, c, cots)
            if d_cams is not None:
                d_cams = d_cams.to(cams.device)
        return (None, d_imgs, d_feat.permute(0, 3, 1, 2) if need_feat else None, d_cams, None, None, None)


# --------------------------------------------------------------------------------------------------
# IBRNet.forward
# --------------------------------------------------------------------------------------------------
class IBRNetAggregate(torch.autograd.Function):
    """raw[R,S,4] = f(rgb_feat[R,S,V,35], ray_diff[R,S,V,4], mask[R,S,V,1]; params blob, pos_enc[S,16])."""

    @staticmethod
    def forward(ctx, rgb_feat, ray_diff, mask, params, pos_enc, anti_alias):
        _lib.require_cuda(rgb_feat, ray_diff, mask, params, pos_enc)
        if rgb_feat.shape[-1] != ROW_CH:
            raise RuntimeError(f'IBRNet kernels are built for {ROW_CH}-channel rows (in_feat_ch=32), got {rgb_feat.shape[-1]}')
        R, S, V = rgb_feat.shape[:3]
        if pos_enc.shape[0] != S:
            raise RuntimeError(f'IBRNet(n_samples={pos_enc.shape[0]}) called with {S} samples per ray '
                               '(mlp_network.py:261 adds pos_encoding[1,n_samples,16])')
        rf, rd, mk = f32c(rgb_feat), f32c(ray_diff), f32c(mask)
        dev = rf.device
        N = R * S
        ps = torch.empty(N, PS_STRIDE, device=dev, dtype=torch.float32)
        raw = torch.empty(R, S, 4, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            st = stream_ptr(dev)
            call('nfb_ibrnet_view_fwd', N, S, V, int(anti_alias), ptr(rf), ptr(rd), ptr(mk), 0, 0, 0, 0,
                 None, None, None, None, None, None, None, ptr(params), ptr(ps), None, _lib.precision_code(), st)
            call('nfb_ibrnet_ray_fwd', R, S, ptr(ps), ptr(params), ptr(pos_enc), ptr(raw), None, None, _lib.precision_code(), st)
        ctx.save_for_backward(rf, rd, mk, params.detach(), pos_enc, ps)
        ctx.dims = (R, S, V, int(anti_alias))
        ctx.precision = _lib.precision_code()
        return raw

    @staticmethod
    def backward(ctx, d_raw):
        rf, rd, mk, params, pos_enc, ps = ctx.saved_tensors
        R, S, V, aa = ctx.dims
        N = R * S
        dev = rf.device
        d_rf = None
        if ctx.needs_input_grad[3]:
            # training: data + parameter gradients from the fp32 recompute kernels
            g = f32c(d_raw)
            d_ps = torch.empty(N, PS_STRIDE, device=dev, dtype=torch.float32)
            d_rf = torch.empty_like(rf)
            d_params = torch.zeros(_lib.PARAM_FLOATS, device=dev, dtype=torch.float32)
            with torch.cuda.device(dev):
                st = stream_ptr(dev)
                call('nfb_ibrnet_ray_wgrad', R, S, ptr(ps), ptr(params), ptr(pos_enc), ptr(g), ptr(d_ps), ptr(d_params), st)
                call('nfb_ibrnet_view_wgrad', N, S, V, aa, ptr(rf), ptr(rd), ptr(mk), 0, 0, 0, 0,
                     None, None, None, None, None, None, None, ptr(params), ptr(ps), ptr(d_ps),
                     ptr(d_rf), None, None, ptr(d_params), st)
            return (d_rf if ctx.needs_input_grad[0] else None), None, None, d_params, None, None
        if ctx.needs_input_grad[0]:
            g = f32c(d_raw)
            d_ps = torch.empty(N, PS_STRIDE, device=dev, dtype=torch.float32)
            d_rf = torch.empty_like(rf)
            with torch.cuda.device(dev):
                st = stream_ptr(dev)
                call('nfb_ibrnet_ray_bwd', R, S, ptr(ps), ptr(params), ptr(pos_enc), ptr(g), ptr(d_ps), None, ctx.precision, st)
                call('nfb_ibrnet_view_bwd', N, S, V, aa, ptr(rf), ptr(rd), ptr(mk), 0, 0, 0, 0,
                     None, None, None, None, None, None, None, ptr(params), ptr(ps), ptr(d_ps),
                     ptr(d_rf), None, None, None, ctx.precision, st)
        return d_rf, None, None, None, None, None


# --------------------------------------------------------------------------------------------------
# raw2outputs
# --------------------------------------------------------------------------------------------------
class Composite(torch.autograd.Function):
    """rgb, depth, weights, alpha, ray_mask = f(raw[R,S,4], z[R,S], pixel_mask[R,S] bool)."""

    @staticmethod
    def forward(ctx, raw, z_vals, pixel_mask, white_bkgd):
        _lib.require_cuda(raw, z_vals, pixel_mask)
        raw_c, z_c = f32c(raw), f32c(z_vals)
        pm = pixel_mask.to(torch.uint8).contiguous()
        R, S = z_c.shape
        dev = raw_c.device
        rgb = torch.empty(R, 3, device=dev, dtype=torch.float32)
        depth = torch.empty(R, device=dev, dtype=torch.float32)
        weights = torch.empty(R, S, device=dev, dtype=torch.float32)
        alpha = torch.empty(R, S, device=dev, dtype=torch.float32)
        ray_mask = torch.empty(R, device=dev, dtype=torch.uint8)
        with torch.cuda.device(dev):
            call('nfb_composite_fwd', R, S, int(white_bkgd), ptr(raw_c), ptr(z_c), ptr(pm), None, 0,
                 ptr(rgb), ptr(depth), ptr(weights), ptr(alpha), ptr(ray_mask), stream_ptr(dev))
        ctx.save_for_backward(raw_c, z_c)
        ctx.white = int(white_bkgd)
        ctx.set_materialize_grads(False)
        ray_mask = ray_mask.view(torch.bool)
        ctx.mark_non_differentiable(ray_mask)
        return rgb, depth, weights, alpha, ray_mask

    @staticmethod
    def backward(ctx, d_rgb, d_depth, d_weights, d_alpha, _d_mask):
        raw_c, z_c = ctx.saved_tensors
        R, S = z_c.shape
        dev = raw_c.device
        d_raw = torch.empty_like(raw_c)
        with torch.cuda.device(dev):
            call('nfb_composite_bwd', R, S, ctx.white, ptr(raw_c), ptr(z_c), ptr(f32c(d_rgb)), ptr(f32c(d_depth)),
                 ptr(f32c(d_weights)), ptr(f32c(d_alpha)), ptr(d_raw), stream_ptr(dev))
        return d_raw, None, None, None


# --------------------------------------------------------------------------------------------------
# non-differentiable helpers
# --------------------------------------------------------------------------------------------------
def coarse_depths(R, S, near, far, inv_uniform, t_rand, device):
    z = torch.empty(R, S, device=device, dtype=torch.float32)
    with torch.cuda.device(device):
        call('nfb_coarse_depths', R, S, float(near), float(far), int(bool(inv_uniform)), ptr(f32c(t_rand)), ptr(z),
             stream_ptr(device))
    return z


def sample_pdf_op(bins, weights, u, want_inds=False):
    _lib.require_cuda(bins, weights, u)
    b, w, uu = f32c(bins), f32c(weights), f32c(u)
    R, M = w.shape
    n = uu.shape[-1]
    u_rows = 1 if uu.dim() == 1 else uu.shape[0]
    samples = torch.empty(R, n, device=b.device, dtype=torch.float32)
    above = torch.empty(R, n, device=b.device, dtype=torch.int64) if want_inds else None
    with torch.cuda.device(b.device):
        call('nfb_sample_pdf', R, M, n, ptr(b), ptr(w), ptr(uu), u_rows, ptr(samples), ptr(above), stream_ptr(b.device))
    return (samples, above) if want_inds else samples


def fine_depths(z_coarse, weights_coarse, u, inv_uniform):
    _lib.require_cuda(z_coarse, weights_coarse, u)
    z, w, uu = f32c(z_coarse), f32c(weights_coarse.detach()), f32c(u)
    R, S = z.shape
    n = uu.shape[-1]
    u_rows = 1 if uu.dim() == 1 else uu.shape[0]
    out = torch.empty(R, S + n, device=z.device, dtype=torch.float32)
    with torch.cuda.device(z.device):
        call('nfb_fine_depths', R, S, n, int(bool(inv_uniform)), ptr(z), ptr(w), ptr(uu), u_rows, ptr(out),
             stream_ptr(z.device))
    return out


# --------------------------------------------------------------------------------------------------
# fused render_rays level: project + gather + IBRNet + composite without materialising [R,S,V,35]
# --------------------------------------------------------------------------------------------------
class RenderLevel(torch.autograd.Function):
    """One level (coarse or fine) of render_rays (render_ray.py:206-213 / 245-253), fused:
    rgb, depth, weights, alpha, ray_mask = f(featmaps[V,32,h,w], imgs[V,H,W,3]; ray_o, ray_d, z, cam,
    params, pos_enc).  Differentiable w.r.t. featmaps and imgs."""

    @staticmethod
    def forward(ctx, featmaps, imgs, ray_o, ray_d, z, cam, params, pos_enc, H, W, anti_alias, white_bkgd, geo_noise=0.0):
        _lib.require_cuda(featmaps, imgs, ray_o, ray_d, z, cam, params, pos_enc)
        feat = channels_last_feat(featmaps)
        imgs_c, o_c, d_c, z_c = f32c(imgs), f32c(ray_o), f32c(ray_d), f32c(z)
        R, S = z_c.shape
        if pos_enc.shape[0] != S:
            raise RuntimeError(f'IBRNet(n_samples={pos_enc.shape[0]}) called with {S} samples per ray')
        V, fh, fw = feat.shape[0], feat.shape[1], feat.shape[2]
        N = R * S
        dev = z_c.device
        ps = torch.empty(N, PS_STRIDE, device=dev, dtype=torch.float32)
        raw = torch.empty(R, S, 4, device=dev, dtype=torch.float32)
        rgb = torch.empty(R, 3, device=dev, dtype=torch.float32)
        depth = torch.empty(R, device=dev, dtype=torch.float32)
        weights = torch.empty(R, S, device=dev, dtype=torch.float32)
        alpha = torch.empty(R, S, device=dev, dtype=torch.float32)
        ray_mask = torch.empty(R, device=dev, dtype=torch.uint8)
        pmask = torch.empty(R, S, device=dev, dtype=torch.uint8)      # per-sample pixel mask, written by the ray stage
        need_params = ctx.needs_input_grad[6]
        need = ctx.needs_input_grad[0] or ctx.needs_input_grad[1] or need_params
        # activation stash for the backward (768 B per (sample, view) row): written only when a data gradient is
        # wanted and no parameter trains (the wgrad kernels recompute the forward)
        use_stash = need and not need_params
        n_stash = _lib.stash_bytes(N, V) if use_stash else 0
        stash = torch.empty(n_stash, device=dev, dtype=torch.uint8) if n_stash else None
        n_rstash = _lib.ray_stash_bytes(R, S) if use_stash else 0
        rstash = torch.empty(n_rstash, device=dev, dtype=torch.uint8) if n_rstash else None
        with torch.cuda.device(dev):
            st = stream_ptr(dev)
            call('nfb_ibrnet_view_fwd', N, S, V, int(anti_alias), None, None, None, H, W, fh, fw,
                 None, ptr(o_c), ptr(d_c), ptr(z_c), ptr(cam), ptr(imgs_c), ptr(feat), ptr(params), ptr(ps),
                 ptr(stash), _lib.precision_code(), st)
            call('nfb_ibrnet_ray_fwd', R, S, ptr(ps), ptr(params), ptr(pos_enc), ptr(raw), ptr(pmask), ptr(rstash), _lib.precision_code(), st)
            if geo_noise:
                # sigma += N(0, geo_noise) (render_ray.py:133-134); additive, so every backward kernel is unchanged (the
                # compositing backward reads this noisy `raw`, the ray stage differentiates its own pre-noise sigma)
                raw[..., 3].add_(torch.randn(R, S, device=dev, dtype=torch.float32), alpha=float(geo_noise))
            call('nfb_composite_fwd', R, S, int(white_bkgd), ptr(raw), ptr(z_c), ptr(pmask), None, 0,
                 ptr(rgb), ptr(depth), ptr(weights), ptr(alpha), ptr(ray_mask), st)
        ctx.stash = stash
        ctx.rstash = rstash
        if need:
            ctx.save_for_backward(feat, imgs_c, o_c, d_c, z_c, cam, params.detach(), pos_enc, ps, raw)
        ctx.dims = (R, S, V, H, W, fh, fw, int(anti_alias), int(white_bkgd))
        ctx.precision = _lib.precision_code()
        ctx.imgs_shape = imgs.shape
        ctx.set_materialize_grads(False)
        ray_mask = ray_mask.view(torch.bool)
        ctx.mark_non_differentiable(ray_mask)
        return rgb, depth, weights, alpha, ray_mask

    @staticmethod
    def backward(ctx, d_rgb, d_depth, d_weights, d_alpha, _d_mask):
        feat, imgs_c, o_c, d_c, z_c, cam, params, pos_enc, ps, raw = ctx.saved_tensors
        R, S, V, H, W, fh, fw, aa, white = ctx.dims
        N = R * S
        dev = z_c.device
        need_feat, need_imgs, need_params = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[6]
        d_params = torch.zeros(_lib.PARAM_FLOATS, device=dev, dtype=torch.float32) if need_params else None
        d_feat = torch.zeros(V, fh, fw, FEAT_CH, device=dev, dtype=torch.float32) if need_feat else None
        d_imgs = torch.zeros(ctx.imgs_shape, device=dev, dtype=torch.float32) if need_imgs else None
        d_raw = torch.empty(R, S, 4, device=dev, dtype=torch.float32)
        d_ps = torch.empty(N, PS_STRIDE, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            st = stream_ptr(dev)
            call('nfb_composite_bwd', R, S, white, ptr(raw), ptr(z_c), ptr(f32c(d_rgb)), ptr(f32c(d_depth)),
                 ptr(f32c(d_weights)), ptr(f32c(d_alpha)), ptr(d_raw), st)
            if need_params:
                call('nfb_ibrnet_ray_wgrad', R, S, ptr(ps), ptr(params), ptr(pos_enc), ptr(d_raw), ptr(d_ps), ptr(d_params), st)
                call('nfb_ibrnet_view_wgrad', N, S, V, aa, None, None, None, H, W, fh, fw,
                     None, ptr(o_c), ptr(d_c), ptr(z_c), ptr(cam), ptr(imgs_c), ptr(feat), ptr(params), ptr(ps),
                     ptr(d_ps), None, ptr(d_feat), ptr(d_imgs), ptr(d_params), st)
            else:
                call('nfb_ibrnet_ray_bwd', R, S, ptr(ps), ptr(params), ptr(pos_enc), ptr(d_raw), ptr(d_ps), ptr(ctx.rstash), ctx.precision, st)
                call('nfb_ibrnet_view_bwd', N, S, V, aa, None, None, None, H, W, fh, fw,
                     None, ptr(o_c), ptr(d_c), ptr(z_c), ptr(cam), ptr(imgs_c), ptr(feat), ptr(params), ptr(ps),
                     ptr(d_ps), None, ptr(d_feat), ptr(d_imgs), ptr(ctx.stash), ctx.precision, st)
        ctx.stash = None
        ctx.rstash = None
        return (d_feat.permute(0, 3, 1, 2) if need_feat else None, d_imgs,
                None, None, None, None, d_params, None, None, None, None, None, None)
