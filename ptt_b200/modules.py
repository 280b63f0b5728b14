"""Registry-compatible nn.Modules for the two hot module classes of PTT.

`PointnetSAModuleVotes` mirrors ptt/models/backbones_3d/pointnet2/pointnet2_modules.py:22-90 and
`TransformerBlock` / `TransformerBlockOffset` mirror ptt/models/transformer_block/variants.py:127-165,
297-334: same constructor keywords, same `forward` signatures and return values, same `state_dict`
keys and shapes (checkpoints load by key and shape, tracker3d_template.py:110-118).

In eval mode the forward runs the fused sm_100a kernels through the C ABI (BatchNorm folded); in
training mode it runs the reference's decomposition -- our CUDA ops under autograd.Functions plus
torch layers -- so BatchNorm statistics and gradients behave as in the reference.  CUDA only.

`register()` installs both classes into the reference's registries (and the `pointnet2_ops._ext`
drop-in) without editing the reference tree.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops


# ------------------------------------------------------------------------------------------------
# autograd wrappers over the `_ext` ops (pointnet2_utils.py:88-122, 214-262)
# ------------------------------------------------------------------------------------------------
class _Gather(torch.autograd.Function):
    @staticmethod
    def forward(ctx, features, idx):
        ctx.save_for_backward(idx)
        ctx.n = features.shape[2]
        return ops.gather_points(features, idx)

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        return ops.gather_points_grad(grad_out.contiguous(), idx, ctx.n), None


class _Group(torch.autograd.Function):
    @staticmethod
    def forward(ctx, features, idx):
        ctx.save_for_backward(idx)
        ctx.n = features.shape[2]
        return ops.group_points(features, idx)

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        return ops.group_points_grad(grad_out.contiguous(), idx, ctx.n), None


# ------------------------------------------------------------------------------------------------
# SharedMLP with the reference's module names (pytorch_utils.py:12-36, 39-91, 158-189)
# ------------------------------------------------------------------------------------------------
class _BN2d(nn.Sequential):
    def __init__(self, c):
        super().__init__()
        self.add_module("bn", nn.BatchNorm2d(c))


class _ConvUnit(nn.Sequential):
    def __init__(self, cin, cout, bn):
        super().__init__()
        conv = nn.Conv2d(cin, cout, kernel_size=(1, 1), bias=not bn)
        nn.init.kaiming_normal_(conv.weight)
        if not bn:
            nn.init.constant_(conv.bias, 0)
        self.add_module("conv", conv)
        if bn:
            self.add_module("normlayer", _BN2d(cout))
        self.add_module("activation", nn.ReLU(inplace=True))


class _SharedMLP(nn.Sequential):
    def __init__(self, spec, bn):
        super().__init__()
        for i in range(len(spec) - 1):
            self.add_module("layer%d" % i, _ConvUnit(spec[i], spec[i + 1], bn))


class PointnetSAModuleVotes(nn.Module):
    def __init__(self, *, mlp, radius=None, nsample=None, bn=True, use_xyz=True, normalize_xyz=False,
                 sample_uniformly=False, sample_method="fps"):
        super().__init__()
        if sample_uniformly:
            raise NotImplementedError("sample_uniformly is not enabled by any PTT config (pointnet2_utils.py:339-348)")
        self.radius = radius
        self.nsample = nsample
        self.use_xyz = use_xyz
        self.normalize_xyz = normalize_xyz
        self.sample_method = sample_method
        mlp_spec = mlp
        if use_xyz and len(mlp_spec) > 0:
            mlp_spec[0] += 3          # in place, like the reference (pointnet2_modules.py:51-53)
        self.mlp_module = _SharedMLP(mlp_spec, bn)
        self._packed = None

    # packed parameters are rebuilt whenever the module may have changed
    def train(self, mode=True):
        self._packed = None
        return super().train(mode)

    def _load_from_state_dict(self, *a, **k):
        self._packed = None
        return super()._load_from_state_dict(*a, **k)

    def _apply(self, fn, *a, **k):
        self._packed = None
        return super()._apply(fn, *a, **k)

    def _pack(self):
        if self._packed is None:
            ws, scales, shifts = [], [], []
            for unit in self.mlp_module:
                w = unit.conv.weight.detach()
                ws.append(w.reshape(w.shape[0], w.shape[1]))
                if hasattr(unit, "normlayer"):
                    bn = unit.normlayer.bn
                    s, t = ops.fold_batchnorm(bn.weight.detach(), bn.bias.detach(), bn.running_mean, bn.running_var, bn.eps)
                else:
                    s, t = None, unit.conv.bias.detach()
                scales.append(s)
                shifts.append(t)
            self._packed = ops.PackedSAMlp(ws, scales, shifts)
        return self._packed

    def _sample(self, xyz, features, npoint):
        if self.sample_method == "fps":
            return ops.furthest_point_sampling(xyz, npoint)
        if self.sample_method in ("rs", "sequence"):
            return torch.arange(npoint, dtype=torch.int32, device=xyz.device).repeat(xyz.size(0), 1)
        if self.sample_method == "ffps":
            f = torch.cat([xyz, features.transpose(1, 2)], dim=2).contiguous()
            d = torch.cdist(f, f).pow(2).contiguous()
            return ops.furthest_point_sampling_with_dist(d, npoint)
        raise NotImplementedError(self.sample_method)

    def forward(self, xyz, features, npoint, inds=None):
        if not xyz.is_cuda:
            raise ops.PttError("ptt_b200.PointnetSAModuleVotes runs on CUDA only (there is no CPU path)")
        xyz = xyz.contiguous()
        if inds is None:
            inds = self._sample(xyz, features, npoint)
        else:
            assert inds.shape[1] == npoint
        inds32 = inds.to(torch.int32).contiguous()
        fused = (not self.training) and self.use_xyz and not (torch.is_grad_enabled() and (
            xyz.requires_grad or (features is not None and features.requires_grad)))
        if fused:
            new_xyz = torch.gather(xyz, 1, inds32.long().unsqueeze(-1).expand(-1, -1, 3))
            idx = ops.ball_query(new_xyz, xyz, self.radius, self.nsample)
            feats_pm = ops.cm_to_pm(features.contiguous()) if features is not None else None
            _, new_features = ops.sa_mlp_fwd(self._pack(), xyz, feats_pm, new_xyz, idx, self.radius, self.normalize_xyz,
                                             want_pm=False, want_cm=True)
            return new_xyz, new_features, inds.to(torch.int64)

        # the reference's decomposition (pointnet2_modules.py:62-90, pointnet2_utils.py:320-380) under autograd
        xyz_flipped = xyz.transpose(1, 2).contiguous()
        new_xyz = _Gather.apply(xyz_flipped, inds32).transpose(1, 2).contiguous()
        idx = ops.ball_query(new_xyz.detach(), xyz.detach(), self.radius, self.nsample)
        grouped_xyz = _Group.apply(xyz_flipped, idx) - new_xyz.transpose(1, 2).unsqueeze(-1)
        if self.normalize_xyz:
            grouped_xyz = grouped_xyz / self.radius
        if features is not None:
            grouped = _Group.apply(features.contiguous(), idx)
            new_features = torch.cat([grouped_xyz, grouped], dim=1) if self.use_xyz else grouped
        else:
            assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
            new_features = grouped_xyz
        new_features = self.mlp_module(new_features)
        new_features = F.max_pool2d(new_features, kernel_size=[1, new_features.size(3)]).squeeze(-1)
        return new_xyz, new_features, inds.to(torch.int64)


class TransformerBlock(nn.Module):
    VARIANT = 0

    def __init__(self, d_points, d_model, k, **kwargs):
        super().__init__()
        self.fc1 = nn.Linear(d_points, d_model)
        self.fc2 = nn.Linear(d_model, d_points)
        self.fc_delta = nn.Sequential(nn.Linear(3, d_model), nn.ReLU(), nn.Linear(d_model, d_model))
        self.fc_gamma = nn.Sequential(nn.Linear(d_model, d_model), nn.ReLU(), nn.Linear(d_model, d_model))
        self.w_qs = nn.Linear(d_model, d_model, bias=False)
        self.w_ks = nn.Linear(d_model, d_model, bias=False)
        self.w_vs = nn.Linear(d_model, d_model, bias=False)
        self.k = k
        self.return_attn = True      # the reference returns (res, attn); its callers use only [0]
        self._packed = None

    def train(self, mode=True):
        self._packed = None
        return super().train(mode)

    def _load_from_state_dict(self, *a, **k):
        self._packed = None
        return super()._load_from_state_dict(*a, **k)

    def _apply(self, fn, *a, **k):
        self._packed = None
        return super()._apply(fn, *a, **k)

    def _pack(self):
        if self._packed is None:
            sd = {k: v.detach() for k, v in self.state_dict().items()}
            self._packed = ops.PackedTransformer(sd, self.k, self.VARIANT)
        return self._packed

    def forward(self, xyz, features):
        if not xyz.is_cuda:
            raise ops.PttError("ptt_b200.TransformerBlock runs on CUDA only (there is no CPU path)")
        xyz = xyz.contiguous()
        features = features.contiguous()
        fused = (not self.training) and not (torch.is_grad_enabled() and (xyz.requires_grad or features.requires_grad))
        if fused:
            r = ops.transformer_block_fwd(self._pack(), xyz, features, want_attn=self.return_attn)
            return r if self.return_attn else (r, None)

        # the reference's decomposition (variants.py:149-165) with the kNN selection on our kernel
        knn_idx = ops.knn(xyz.detach(), self.k).long()
        B, n, _ = xyz.shape

        def take(points):
            flat = knn_idx.reshape(B, n * self.k, 1).expand(-1, -1, points.shape[-1])
            return torch.gather(points, 1, flat).reshape(B, n, self.k, -1)

        pre = features
        x = self.fc1(features)
        q, kk, v = self.w_qs(x), take(self.w_ks(x)), take(self.w_vs(x))
        pos_enc = self.fc_delta(xyz[:, :, None] - take(xyz))
        attn = self.fc_gamma(q[:, :, None] - kk + pos_enc)
        attn = F.softmax(attn / math.sqrt(kk.size(-1)), dim=-2)
        res = torch.einsum("bmnf,bmnf->bmf", attn, v + pos_enc)
        if self.VARIANT == 1:
            res = x - res
        return self.fc2(res) + pre, attn


class TransformerBlockOffset(TransformerBlock):
    VARIANT = 1


class TransformerBlockSTD(nn.Module):
    """Dense n x n dot-product attention (variants.py:12-40); same constructor, forward and state_dict."""

    def __init__(self, d_points, d_model, k, **kwargs):
        super().__init__()
        self.fc1 = nn.Linear(d_points, d_model)
        self.fc2 = nn.Linear(d_model, d_points)
        self.fc_delta = nn.Sequential(nn.Linear(3, d_model), nn.ReLU(), nn.Linear(d_model, d_model))
        self.w_qs = nn.Linear(d_model, d_model, bias=False)
        self.w_ks = nn.Linear(d_model, d_model, bias=False)
        self.w_vs = nn.Linear(d_model, d_model, bias=False)
        self.k = k
        self._packed = None

    def train(self, mode=True):
        self._packed = None
        return super().train(mode)

    def _load_from_state_dict(self, *a, **k):
        self._packed = None
        return super()._load_from_state_dict(*a, **k)

    def _apply(self, fn, *a, **k):
        self._packed = None
        return super()._apply(fn, *a, **k)

    def forward(self, xyz, features):
        if not xyz.is_cuda:
            raise ops.PttError("ptt_b200.TransformerBlockSTD runs on CUDA only (there is no CPU path)")
        xyz, features = xyz.contiguous(), features.contiguous()
        fused = (not self.training) and not (torch.is_grad_enabled() and (xyz.requires_grad or features.requires_grad))
        if fused:
            if self._packed is None:
                self._packed = ops.PackedTransformerSTD({k: v.detach() for k, v in self.state_dict().items()})
            return ops.transformer_std_fwd(self._packed, xyz, features, want_attn=True)
        pre = features
        x = self.fc1(features)
        q, k, v = self.w_qs(x), self.w_ks(x), self.w_vs(x)
        attn = F.softmax(q @ k.transpose(1, 2) / math.sqrt(k.size(-1)), dim=-1)
        res = attn @ (v + self.fc_delta(xyz))
        return self.fc2(res) + pre, attn


def register(install_ext=True):
    """Install the B200 modules into the reference's registries (the reference must be importable
    as `ptt`):  pointnet2_modules.PointnetSAModuleVotes (looked up at construction time by
    pointnet2_backbone.py:22 and box_voting_head.py:19) and transformer_block.__all__ (:7-17).
    With install_ext, `pointnet2_ops._ext` is provided first (ptt_b200.install_dropin)."""
    if install_ext:
        from . import install_dropin
        install_dropin()
    from ptt.models.backbones_3d.pointnet2 import pointnet2_modules
    from ptt.models import transformer_block

    pointnet2_modules.PointnetSAModuleVotes = PointnetSAModuleVotes
    transformer_block.__all__["TransformerBlock"] = TransformerBlock
    transformer_block.__all__["TransformerBlockOffset"] = TransformerBlockOffset
    transformer_block.__all__["TransformerBlockSTD"] = TransformerBlockSTD
    return pointnet2_modules, transformer_block
