"""CPU oracle for the GNT forward (SURVEY.md 8 row a15 / BASELINE config 5).

TEST INFRASTRUCTURE ONLY (same rules as ``ibrnet_oracle.py``): imported by ``tests/`` and by ``bench.py``'s
CPU-baseline leg, never by ``nerfool_b200/``.

A functional (no ``nn.Module``) restatement in plain PyTorch ops of

    Embedder                     /root/reference/gnt/transformer_network.py:6-37
    FeedForward                  :40-52
    Attention2D                  :55-89     (view transformer: subtraction attention, softmax over views per channel)
    Transformer2D                :93-113
    Attention ("qk") / Transformer   :121-202   (ray transformer, 4 heads, no mask)
    GNT.forward                  :270-309
    render_rays (coarse level)   /root/reference/gnt/render_ray.py:196-250

with dropout as the identity (``.eval()``).  ``p`` is a mapping keyed by the reference's ``state_dict`` names.
Parity pinning: ``oracle/make_golden_gnt.py`` runs the unmodified reference module on seeded inputs and commits
``tests/golden/gnt_d2.npz`` (state dict, inputs, outputs with ``ret_alpha``); ``tests/test_oracle_golden.py`` holds
this file to those vectors.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def embed(x: torch.Tensor, num_freqs: int = 10) -> torch.Tensor:
    """Embedder(include_input, log_sampling, max_freq_log2=9, num_freqs=10, [sin, cos]) (:6-37): 3 -> 63."""
    freqs = 2.0 ** torch.linspace(0.0, float(num_freqs - 1), steps=num_freqs)
    out = [x]
    for f in freqs:
        out.append(torch.sin(x * f))
        out.append(torch.cos(x * f))
    return torch.cat(out, -1)


def _lin(p, name, x):
    return F.linear(x, p[name + '.weight'], p.get(name + '.bias'))


def _ln(p, name, x, eps):
    return F.layer_norm(x, (x.shape[-1],), p[name + '.weight'], p[name + '.bias'], eps)


def _ff(p, pre, x):
    return _lin(p, pre + '.fc2', F.relu(_lin(p, pre + '.fc1', x)))


def view_attention(p, pre, q, k, pos, mask):
    """Attention2D.forward (:74-89).  q [R,S,D], k [R,S,V,D], pos [R,S,V,4], mask [R,S,V,1]."""
    q = _lin(p, pre + '.q_fc', q)
    k = _lin(p, pre + '.k_fc', k)
    v = _lin(p, pre + '.v_fc', k)                      # applied to the projected k, as the reference does
    pos = _lin(p, pre + '.pos_fc.2', F.relu(_lin(p, pre + '.pos_fc.0', pos)))
    attn = k - q[:, :, None, :] + pos
    attn = _lin(p, pre + '.attn_fc.2', F.relu(_lin(p, pre + '.attn_fc.0', attn)))
    attn = attn.masked_fill(mask == 0, -1e9)
    attn = torch.softmax(attn, dim=-2)
    x = ((v + pos) * attn).sum(dim=2)
    return _lin(p, pre + '.out_fc', x)


def view_transformer(p, pre, q, k, pos, mask):
    """Transformer2D.forward (:102-113), LayerNorm eps 1e-6."""
    x = view_attention(p, pre + '.attn', _ln(p, pre + '.attn_norm', q, 1e-6), k, pos, mask) + q
    return _ff(p, pre + '.ff', _ln(p, pre + '.ff_norm', x, 1e-6)) + x


def ray_attention(p, pre, x, n_heads=4):
    """Attention.forward, mode "qk" (:141-171).  x [R,S,D] -> (out [R,S,D], attn [R,H,S,S])."""
    R, S, Dm = x.shape
    q = _lin(p, pre + '.q_fc', x).view(R, S, n_heads, -1).permute(0, 2, 1, 3)
    k = _lin(p, pre + '.k_fc', x).view(R, S, n_heads, -1).permute(0, 2, 1, 3)
    v = _lin(p, pre + '.v_fc', x).view(R, S, n_heads, -1).permute(0, 2, 1, 3)
    attn = torch.softmax(torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(q.shape[-1]), dim=-1)
    out = torch.matmul(attn, v).permute(0, 2, 1, 3).contiguous().view(R, S, -1)
    return _lin(p, pre + '.out_fc', out), attn


def ray_transformer(p, pre, x):
    """Transformer.forward (:185-202) -> (x, attn.mean(heads)[:, 0])."""
    a, attn = ray_attention(p, pre + '.attn', _ln(p, pre + '.attn_norm', x, 1e-6))
    x = a + x
    x = _ff(p, pre + '.ff', _ln(p, pre + '.ff_norm', x, 1e-6)) + x
    return x, attn.mean(dim=1)[:, 0]


def gnt_forward(p, depth, rgb_feat, ray_diff, mask, pts, ray_d, ret_alpha=True):
    """GNT.forward (:270-309).  rgb_feat [R,S,V,35], ray_diff [R,S,V,4], mask [R,S,V,1], pts [R,S,3], ray_d [R,3]
    -> [R,3] or [R,3+S]."""
    viewdirs = ray_d / torch.norm(ray_d, dim=-1, keepdim=True)
    viewdirs = embed(viewdirs.reshape(-1, 3).to(pts.dtype))
    pts_ = embed(pts.reshape(-1, 3)).reshape(*pts.shape[:-1], -1)
    viewdirs_ = viewdirs[:, None].expand(pts_.shape)
    feat = _lin(p, 'rgbfeat_fc.2', F.relu(_lin(p, 'rgbfeat_fc.0', rgb_feat)))
    q = feat.max(dim=2)[0]
    attn = None
    for i in range(depth):
        q = view_transformer(p, f'view_crosstrans.{i}', q, feat, ray_diff, mask)
        if i % 2 == 0:
            q = torch.cat((q, pts_, viewdirs_), dim=-1)
            q = _lin(p, f'q_fcs.{i}.2', F.relu(_lin(p, f'q_fcs.{i}.0', q)))
        q, attn = ray_transformer(p, f'view_selftrans.{i}', q)
    h = _ln(p, 'norm', q, 1e-5)
    out = _lin(p, 'rgb_fc', h.mean(dim=1))
    return torch.cat([out, attn], dim=1) if ret_alpha else out


def random_gnt_params(depth: int, seed: int, width: int = 64):
    """Random parameters with the reference's names / shapes (PyTorch-default-like uniform init; biases non-zero)."""
    g = torch.Generator().manual_seed(seed)
    p = {}

    def lin(name, n_out, n_in, bias=True):
        b = 1.0 / math.sqrt(n_in)
        p[name + '.weight'] = (torch.rand(n_out, n_in, generator=g) * 2 - 1) * b
        if bias:
            p[name + '.bias'] = (torch.rand(n_out, generator=g) * 2 - 1) * b

    def ln(name):
        p[name + '.weight'] = 1.0 + 0.1 * torch.randn(width, generator=g)
        p[name + '.bias'] = 0.1 * torch.randn(width, generator=g)

    lin('rgbfeat_fc.0', width, 35); lin('rgbfeat_fc.2', width, width)
    for i in range(depth):
        for kind, pre in (('ray', f'view_selftrans.{i}'), ('view', f'view_crosstrans.{i}')):
            ln(pre + '.attn_norm'); ln(pre + '.ff_norm')
            lin(pre + '.ff.fc1', 4 * width, width); lin(pre + '.ff.fc2', width, 4 * width)
            for n in ('q_fc', 'k_fc', 'v_fc'):
                lin(pre + '.attn.' + n, width, width, bias=False)
            if kind == 'view':
                lin(pre + '.attn.pos_fc.0', width // 8, 4); lin(pre + '.attn.pos_fc.2', width, width // 8)
                lin(pre + '.attn.attn_fc.0', width // 8, width); lin(pre + '.attn.attn_fc.2', width, width // 8)
            lin(pre + '.attn.out_fc', width, width)
        if i % 2 == 0:
            lin(f'q_fcs.{i}.0', width, width + 126); lin(f'q_fcs.{i}.2', width, width)
    ln('norm')
    lin('rgb_fc', 3, width)
    return p
