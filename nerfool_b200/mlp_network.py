"""Drop-in for ``ibrnet.mlp_network.IBRNet`` (/root/reference/ibrnet/mlp_network.py:152-274).

An ``nn.Module`` with the reference's parameter / buffer names and shapes (so checkpoints load, DDP /
DataParallel wrapping and ``.to()`` work) whose ``forward`` runs the CUDA view-stage + ray-stage kernels.
The sub-modules exist only as parameter containers; their ``forward`` is never called.  Gradients: data gradients
(feature maps / source images, what the PGD attack optimises) always; parameter gradients (training) whenever a
parameter has ``requires_grad`` -- then the backward runs the wgrad kernels (nfb_ibrnet_{view,ray}_wgrad)."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import ops

# order of the tensors inside the parameter blob (include/nerfool_b200.h)
PARAM_ORDER = [
    's',
    'ray_dir_fc.0.weight', 'ray_dir_fc.0.bias', 'ray_dir_fc.2.weight', 'ray_dir_fc.2.bias',
    'base_fc.0.weight', 'base_fc.0.bias', 'base_fc.2.weight', 'base_fc.2.bias',
    'vis_fc.0.weight', 'vis_fc.0.bias', 'vis_fc.2.weight', 'vis_fc.2.bias',
    'vis_fc2.0.weight', 'vis_fc2.0.bias', 'vis_fc2.2.weight', 'vis_fc2.2.bias',
    'geometry_fc.0.weight', 'geometry_fc.0.bias', 'geometry_fc.2.weight', 'geometry_fc.2.bias',
    'ray_attention.w_qs.weight', 'ray_attention.w_ks.weight', 'ray_attention.w_vs.weight',
    'ray_attention.fc.weight', 'ray_attention.layer_norm.weight', 'ray_attention.layer_norm.bias',
    'out_geometry_fc.0.weight', 'out_geometry_fc.0.bias', 'out_geometry_fc.2.weight', 'out_geometry_fc.2.bias',
    'rgb_fc.0.weight', 'rgb_fc.0.bias', 'rgb_fc.2.weight', 'rgb_fc.2.bias', 'rgb_fc.4.weight', 'rgb_fc.4.bias',
]
PARAM_FLOATS = 20136


def pack_params(tensors: dict, device=None) -> torch.Tensor:
    """state_dict-like mapping -> flat fp32 blob in PARAM_ORDER ('s' defaults to 0 when absent)."""
    parts = []
    for name in PARAM_ORDER:
        if name == 's' and name not in tensors:
            parts.append(torch.zeros(1, device=device))
            continue
        parts.append(tensors[name].detach().reshape(-1).float())
    blob = torch.cat([p.to(device) if device is not None else p for p in parts])
    assert blob.numel() == PARAM_FLOATS, blob.numel()
    return blob.contiguous()


def _weights_init(m):
    # default tensorflow initialisation of linear layers (mlp_network.py:137-141)
    if isinstance(m, nn.Linear):
        nn.init.kaiming_normal_(m.weight.data)
        if m.bias is not None:
            nn.init.zeros_(m.bias.data)


class _RayAttentionParams(nn.Module):
    """Parameter container named like MultiHeadAttention(4, 16, 4, 4) (mlp_network.py:69-88)."""

    def __init__(self, n_head=4, d_model=16, d_k=4, d_v=4):
        super().__init__()
        self.n_head, self.d_k, self.d_v = n_head, d_k, d_v
        self.w_qs = nn.Linear(d_model, n_head * d_k, bias=False)
        self.w_ks = nn.Linear(d_model, n_head * d_k, bias=False)
        self.w_vs = nn.Linear(d_model, n_head * d_v, bias=False)
        self.fc = nn.Linear(n_head * d_v, d_model, bias=False)
        self.layer_norm = nn.LayerNorm(d_model, eps=1e-6)


class IBRNet(nn.Module):
    def __init__(self, args, in_feat_ch=32, n_samples=64, **kwargs):
        super().__init__()
        if in_feat_ch != 32:
            raise NotImplementedError('nerfool_b200 IBRNet kernels are built for in_feat_ch=32 '
                                      '(coarse_feat_dim / fine_feat_dim of every shipped config)')
        self.args = args
        self.anti_alias_pooling = args.anti_alias_pooling
        if self.anti_alias_pooling:
            self.s = nn.Parameter(torch.tensor(0.2), requires_grad=True)
        act = nn.ELU(inplace=True)
        self.n_samples = n_samples
        # construction order == the reference's, so the same torch seed gives the same initial weights
        self.ray_dir_fc = nn.Sequential(nn.Linear(4, 16), act, nn.Linear(16, in_feat_ch + 3), act)
        self.base_fc = nn.Sequential(nn.Linear((in_feat_ch + 3) * 3, 64), act, nn.Linear(64, 32), act)
        self.vis_fc = nn.Sequential(nn.Linear(32, 32), act, nn.Linear(32, 33), act)
        self.vis_fc2 = nn.Sequential(nn.Linear(32, 32), act, nn.Linear(32, 1), nn.Sigmoid())
        self.geometry_fc = nn.Sequential(nn.Linear(32 * 2 + 1, 64), act, nn.Linear(64, 16), act)
        self.ray_attention = _RayAttentionParams(4, 16, 4, 4)
        self.out_geometry_fc = nn.Sequential(nn.Linear(16, 16), act, nn.Linear(16, 1), nn.ReLU())
        self.rgb_fc = nn.Sequential(nn.Linear(32 + 1 + 4, 16), act, nn.Linear(16, 8), act, nn.Linear(8, 1))
        self.register_buffer('pos_encoding', self.posenc(d_hid=16, n_samples=self.n_samples))
        self.base_fc.apply(_weights_init)
        self.vis_fc2.apply(_weights_init)
        self.vis_fc.apply(_weights_init)
        self.geometry_fc.apply(_weights_init)
        self.rgb_fc.apply(_weights_init)
        self._blob_key = None
        self._blob = None

    def posenc(self, d_hid, n_samples):
        # sinusoid table (mlp_network.py:210-220), float64 numpy then cast
        pos = np.arange(n_samples, dtype=np.float64)[:, None]
        j = np.arange(d_hid)
        table = pos / np.power(10000, 2 * (j // 2) / d_hid)[None, :]
        table[:, 0::2] = np.sin(table[:, 0::2])
        table[:, 1::2] = np.cos(table[:, 1::2])
        return torch.from_numpy(table).float().unsqueeze(0)

    # ------------------------------------------------------------------------------------------
    def param_blob(self) -> torch.Tensor:
        """Flat fp32 copy of the parameters in the C-ABI order.  Inference / attack (``.eval()`` -- the attack
        drivers call model.switch_to_eval(), eval_adv.py:541 --, or no parameter requires a gradient, or grad mode
        is off): a detached blob, rebuilt only when a parameter changed; the backward then runs the fast
        data-gradient kernels only.  (The reference's eager ``loss.backward()`` would also fill the parameters'
        ``.grad`` in eval mode; nothing in the attack reads them, so they are not produced.)  Training
        (``.train()``, train.py:317-327): a differentiable ``torch.cat`` of the parameters, so the gradient blob
        the wgrad kernels return is split back onto ``.grad`` of every tensor by autograd."""
        sd = dict(self.named_parameters())
        if self.training and torch.is_grad_enabled() and any(p.requires_grad for p in sd.values()):
            dev = next(self.parameters()).device
            parts = []
            for name in PARAM_ORDER:
                if name == 's' and name not in sd:
                    parts.append(torch.zeros(1, device=dev))
                else:
                    parts.append(sd[name].reshape(-1).float())
            return torch.cat(parts)
        key = tuple((p.data_ptr(), p._version) for p in sd.values())
        if self._blob is None or key != self._blob_key:
            dev = next(self.parameters()).device
            self._blob = pack_params(sd, device=dev)
            self._blob_key = key
        return self._blob

    def forward(self, rgb_feat, ray_diff, mask):
        """
        :param rgb_feat: rgbs and image features [n_rays, n_samples, n_views, n_feat]
        :param ray_diff: ray direction difference [n_rays, n_samples, n_views, 4]
        :param mask: [n_rays, n_samples, n_views, 1]
        :return: rgb and density output, [n_rays, n_samples, 4]
        """
        return ops.IBRNetAggregate.apply(rgb_feat, ray_diff, mask, self.param_blob(), self.pos_encoding[0],
                                         bool(self.anti_alias_pooling))
