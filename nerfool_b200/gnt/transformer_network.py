"""Drop-in for ``gnt.transformer_network.GNT`` (/root/reference/gnt/transformer_network.py:205-309).

An ``nn.Module`` with the reference's parameter names, shapes and construction order (checkpoints load strictly, the
same torch seed gives the same initial weights) whose ``forward`` runs ``nfb_gnt_fwd``.  The sub-modules are parameter
containers; their ``forward`` is never called.  Differentiable w.r.t. its sampled inputs ``rgb_feat`` and ``ray_diff``
(``nfb_gnt_bwd``: what eval/gnt/eval_adv.py:282-545 back-propagates to the perturbation and, with ``--perturb_camera``, to the
source poses); parameter gradients (training) are not built and asking for them raises (no silent fallback)."""
from __future__ import annotations

import ctypes

import torch
import torch.nn as nn

from .. import _lib
from .._lib import call, f32c, ptr, stream_ptr


class _FeedForward(nn.Module):          # transformer_network.py:40-46
    def __init__(self, dim, hid_dim):
        super().__init__()
        self.fc1 = nn.Linear(dim, hid_dim)
        self.fc2 = nn.Linear(hid_dim, dim)


class _Attention2D(nn.Module):          # :55-72
    def __init__(self, dim):
        super().__init__()
        self.q_fc = nn.Linear(dim, dim, bias=False)
        self.k_fc = nn.Linear(dim, dim, bias=False)
        self.v_fc = nn.Linear(dim, dim, bias=False)
        self.pos_fc = nn.Sequential(nn.Linear(4, dim // 8), nn.ReLU(), nn.Linear(dim // 8, dim))
        self.attn_fc = nn.Sequential(nn.Linear(dim, dim // 8), nn.ReLU(), nn.Linear(dim // 8, dim))
        self.out_fc = nn.Linear(dim, dim)


class _Attention(nn.Module):            # :121-139, attn_mode "qk"
    def __init__(self, dim, n_heads):
        super().__init__()
        self.q_fc = nn.Linear(dim, dim, bias=False)
        self.k_fc = nn.Linear(dim, dim, bias=False)
        self.v_fc = nn.Linear(dim, dim, bias=False)
        self.out_fc = nn.Linear(dim, dim)
        self.n_heads = n_heads


class _TransformerParams(nn.Module):    # Transformer2D :93-100 / Transformer :175-183 (same member order)
    def __init__(self, dim, make_attn):
        super().__init__()
        self.attn_norm = nn.LayerNorm(dim, eps=1e-6)
        self.ff_norm = nn.LayerNorm(dim, eps=1e-6)
        self.ff = _FeedForward(dim, 4 * dim)
        self.attn = make_attn()          # created after ff, as in the reference (RNG consumption order)


# (block-relative name inside the C-ABI layer block, state_dict name pattern)
_VIEW = ['attn_norm.weight', 'attn_norm.bias', 'attn.q_fc.weight', 'attn.k_fc.weight', 'attn.v_fc.weight',
         'attn.pos_fc.0.weight', 'attn.pos_fc.0.bias', 'attn.pos_fc.2.weight', 'attn.pos_fc.2.bias',
         'attn.attn_fc.0.weight', 'attn.attn_fc.0.bias', 'attn.attn_fc.2.weight', 'attn.attn_fc.2.bias',
         'attn.out_fc.weight', 'attn.out_fc.bias', 'ff_norm.weight', 'ff_norm.bias',
         'ff.fc1.weight', 'ff.fc1.bias', 'ff.fc2.weight', 'ff.fc2.bias']
_QFC = ['0.weight', '0.bias', '2.weight', '2.bias']
_RAY = ['attn_norm.weight', 'attn_norm.bias', 'attn.q_fc.weight', 'attn.k_fc.weight', 'attn.v_fc.weight',
        'attn.out_fc.weight', 'attn.out_fc.bias', 'ff_norm.weight', 'ff_norm.bias',
        'ff.fc1.weight', 'ff.fc1.bias', 'ff.fc2.weight', 'ff.fc2.bias']
_HEAD = ['rgbfeat_fc.0.weight', 'rgbfeat_fc.0.bias', 'rgbfeat_fc.2.weight', 'rgbfeat_fc.2.bias']
_TAIL = ['norm.weight', 'norm.bias', 'rgb_fc.weight', 'rgb_fc.bias']


def blob_layout(depth: int):
    """[(state_dict name, offset in floats)] of every tensor in the nfb_gnt_fwd parameter blob (from the library)."""
    lib = _lib.load()
    off = lambda n: int(lib.nfb_gnt_param_offset(depth, n.encode()))     # noqa: E731
    out = [(n, off(n)) for n in _HEAD]
    base, size = off('layer0'), off('layer_size')
    for i in range(depth):
        out += [(f'view_crosstrans.{i}.{n}', base + i * size + off('view.' + n)) for n in _VIEW]
        if i % 2 == 0:
            out += [(f'q_fcs.{i}.{n}', base + i * size + off('q_fc.' + n)) for n in _QFC]
        out += [(f'view_selftrans.{i}.{n}', base + i * size + off('ray.' + n)) for n in _RAY]
    out += [(n, off(n)) for n in _TAIL]
    assert all(o >= 0 for _, o in out)
    return out


def pack_params(tensors: dict, depth: int, device=None) -> torch.Tensor:
    """state_dict-like mapping -> flat fp32 blob in the C-ABI layout (unused q_fc slots of odd layers stay zero)."""
    n = int(_lib.load().nfb_gnt_param_floats(depth))
    blob = torch.zeros(n, dtype=torch.float32, device=device)
    for name, o in blob_layout(depth):
        t = tensors[name].detach().reshape(-1).float()
        blob[o:o + t.numel()] = t.to(blob.device)
    return blob


# rays per nfb_gnt_bwd call: the backward's workspace (checkpoints + per-row buffers) is ~ 0.9 MB per ray at S = 64, V = 8, depth 4
BWD_WORKSPACE_GIB = float(__import__('os').environ.get('NFB_GNT_BWD_WS_GIB', '24'))


class _GNTFn(torch.autograd.Function):
    """out = GNT(rgb_feat, ray_diff, mask, pts, ray_d).  Without a gradient request: nfb_gnt_fwd (tensor-core forward).  With one, and if the
    backward's workspace for all rays fits the cap: nfb_gnt_fwd_save (the fp32 checkpointing forward IS the forward) and nfb_gnt_bwd_saved
    (reverse sweep on the kept workspace); otherwise nfb_gnt_fwd now and nfb_gnt_bwd (forward re-run + sweep) in ray chunks later."""

    @staticmethod
    def forward(ctx, rgb_feat, ray_diff, mask, pts, ray_d, blob, depth, ret_alpha):
        rf, rd, mk, pt, dd = f32c(rgb_feat.detach()), f32c(ray_diff.detach()), f32c(mask.detach()), f32c(pts.detach()), f32c(ray_d.detach())
        R, S, V = rf.shape[:3]
        dev = rf.device
        lib = _lib.load()
        out = torch.empty(R, 3 + S if ret_alpha else 3, device=dev, dtype=torch.float32)
        ctx.ws = None
        want_grad = rgb_feat.requires_grad or ray_diff.requires_grad
        bwd_bytes = int(lib.nfb_gnt_bwd_workspace_bytes(R, S, V, depth)) if (want_grad and R > 0) else 0
        with torch.cuda.device(dev):
            if want_grad and 0 < bwd_bytes <= BWD_WORKSPACE_GIB * 2 ** 30 and R * S * V < 2 ** 31:
                ws = torch.empty(bwd_bytes, device=dev, dtype=torch.uint8)
                call('nfb_gnt_fwd_save', R, S, V, depth, int(bool(ret_alpha)), ptr(rf), ptr(rd), ptr(mk), ptr(pt), ptr(dd),
                     ptr(blob), ptr(out), ptr(ws), ctypes.c_size_t(bwd_bytes), stream_ptr(dev))
                ctx.ws = ws
            else:
                nbytes = int(lib.nfb_gnt_workspace_bytes(R, S, V))
                ws = torch.empty(max(nbytes, 16), device=dev, dtype=torch.uint8)
                call('nfb_gnt_fwd', R, S, V, depth, int(bool(ret_alpha)), ptr(rf), ptr(rd), ptr(mk), ptr(pt), ptr(dd),
                     ptr(blob), ptr(out), ptr(ws), ctypes.c_size_t(nbytes), _lib.precision_code(), stream_ptr(dev))
        ctx.save_for_backward(rf, rd, mk, pt, dd, blob)
        ctx.meta = (depth, bool(ret_alpha), ray_diff.requires_grad)
        return out

    @staticmethod
    def backward(ctx, d_out):
        rf, rd, mk, pt, dd, blob = ctx.saved_tensors
        depth, ret_alpha, want_rd = ctx.meta
        R, S, V = rf.shape[:3]
        dev = rf.device
        g = f32c(d_out)
        d_rf = torch.empty_like(rf)
        d_rd = torch.empty_like(rd) if want_rd else None
        lib = _lib.load()
        ws, ctx.ws = ctx.ws, None            # the sweep consumes the saved workspace: a second backward re-runs the forward
        with torch.cuda.device(dev):
            if ws is not None:
                call('nfb_gnt_bwd_saved', R, S, V, depth, int(ret_alpha), ptr(rf), ptr(rd), ptr(mk), ptr(pt), ptr(dd), ptr(blob), ptr(g),
                     ptr(d_rf), ptr(d_rd) if want_rd else None, ptr(ws), ctypes.c_size_t(ws.numel()), stream_ptr(dev))
                return d_rf, d_rd, None, None, None, None, None, None
            per_ray = max(int(lib.nfb_gnt_bwd_workspace_bytes(1, S, V, depth)), 1)
            chunk = max(1, min(R, int(BWD_WORKSPACE_GIB * 2 ** 30) // per_ray, (2 ** 31 - 1) // (S * V)))
            nbytes = int(lib.nfb_gnt_bwd_workspace_bytes(min(chunk, max(R, 1)), S, V, depth))
            ws = torch.empty(max(nbytes, 16), device=dev, dtype=torch.uint8)
            for r0 in range(0, R, chunk):
                r1 = min(R, r0 + chunk)
                call('nfb_gnt_bwd', r1 - r0, S, V, depth, int(ret_alpha), ptr(rf[r0:r1]), ptr(rd[r0:r1]), ptr(mk[r0:r1]), ptr(pt[r0:r1]),
                     ptr(dd[r0:r1]), ptr(blob), ptr(g[r0:r1]), ptr(d_rf[r0:r1]), ptr(d_rd[r0:r1]) if want_rd else None,
                     ptr(ws), ctypes.c_size_t(nbytes), stream_ptr(dev))
        return d_rf, d_rd, None, None, None, None, None, None


class GNT(nn.Module):
    def __init__(self, args, in_feat_ch=32, posenc_dim=3, viewenc_dim=3, ret_alpha=False):
        super().__init__()
        if in_feat_ch != 32 or args.netwidth != 64:
            raise NotImplementedError('nerfool_b200 GNT kernels are built for in_feat_ch=32, netwidth=64 '
                                      '(every shipped GNT config: eval/gnt/config.py:111, configs/gnt/*.txt)')
        if posenc_dim != 63 or viewenc_dim != 63:
            raise NotImplementedError('nerfool_b200 GNT kernels are built for the 63-d positional encodings the '
                                      'reference constructs the model with (gnt/model.py:25-26)')
        w = args.netwidth
        self.rgbfeat_fc = nn.Sequential(nn.Linear(in_feat_ch + 3, w), nn.ReLU(), nn.Linear(w, w))
        # construction order == the reference's (:215-246), so the same torch seed gives the same initial weights
        self.view_selftrans = nn.ModuleList([])
        self.view_crosstrans = nn.ModuleList([])
        self.q_fcs = nn.ModuleList([])
        for i in range(args.trans_depth):
            self.view_crosstrans.append(_TransformerParams(w, lambda: _Attention2D(w)))
            self.view_selftrans.append(_TransformerParams(w, lambda: _Attention(w, 4)))
            if i % 2 == 0:
                self.q_fcs.append(nn.Sequential(nn.Linear(w + posenc_dim + viewenc_dim, w), nn.ReLU(), nn.Linear(w, w)))
            else:
                self.q_fcs.append(nn.Identity())
        self.posenc_dim = posenc_dim
        self.viewenc_dim = viewenc_dim
        self.ret_alpha = ret_alpha
        self.depth = args.trans_depth
        self.norm = nn.LayerNorm(w)
        self.rgb_fc = nn.Linear(w, 3)
        self._blob = None
        self._blob_key = None

    def param_blob(self) -> torch.Tensor:
        sd = dict(self.named_parameters())
        key = tuple((p.data_ptr(), p._version) for p in sd.values())
        if self._blob is None or key != self._blob_key:
            self._blob = pack_params(sd, self.depth, device=next(self.parameters()).device)
            self._blob_key = key
        return self._blob

    def forward(self, rgb_feat, ray_diff, mask, pts, ray_d):
        """
        :param rgb_feat: [n_rays, n_samples, n_views, 35]   :param ray_diff: [.., n_views, 4]   :param mask: [.., n_views, 1]
        :param pts: [n_rays, n_samples, 3]                  :param ray_d: [n_rays, 3]
        :return: [n_rays, 3] or, with ret_alpha, [n_rays, 3 + n_samples]
        """
        _lib.require_cuda(rgb_feat, ray_diff, mask, pts, ray_d)
        if torch.is_grad_enabled() and self.training and any(p.requires_grad for p in self.parameters()):
            raise NotImplementedError('nerfool_b200.gnt.GNT: parameter gradients (training) are not built; the data gradient '
                                      '(rgb_feat, ray_diff) is -- freeze the parameters or call .eval()')
        if self.training:
            raise NotImplementedError('nerfool_b200.gnt.GNT runs in eval mode only (dropout is the identity); call .eval()')
        if rgb_feat.shape[-1] != 35:
            raise RuntimeError(f'GNT kernels are built for 35-channel rows, got {rgb_feat.shape[-1]}')
        return _GNTFn.apply(rgb_feat, ray_diff, mask, pts, ray_d, self.param_blob(), self.depth, bool(self.ret_alpha))
