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


# ------------------------------------------------------------------------------------------------
# Packed-parameter cache and the fused / autograd switch shared by every module below
# ------------------------------------------------------------------------------------------------
class _PackCache:
    """Packed parameter images are rebuilt whenever the parameters may have changed: on train() / _apply() /
    load_state_dict(), and whenever the identity or the in-place version counter of any parameter or buffer differs
    from what it was at pack time (optimizer.step() with frozen BatchNorm, `p.mul_()`, `load_state_dict`, a moved
    module).  Writes that bypass the version counter (`p.data.copy_(..)`, raw-pointer writes) need `invalidate()`."""

    _packed = None
    _packed_key = None

    def invalidate(self):
        self._packed = None
        self._packed_key = None

    def _param_key(self):
        return tuple((t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers()))

    def _cached(self, build):
        key = self._param_key()
        if self._packed is None or self._packed_key != key:
            self._packed = build()
            self._packed_key = key
        return self._packed

    def train(self, mode=True):
        self.invalidate()
        return super().train(mode)

    def _load_from_state_dict(self, *a, **k):
        self.invalidate()
        return super()._load_from_state_dict(*a, **k)

    def _apply(self, fn, *a, **k):
        self.invalidate()
        return super()._apply(fn, *a, **k)

    def _use_fused(self, *tensors):
        """The fused C path returns tensors without an autograd graph, so it is taken only when nobody can ask for
        gradients: eval mode, and either grad mode off (the reference's evaluation loop runs under no_grad) or neither
        an input nor a parameter of this module requires grad.  eval() with grad enabled (frozen-BatchNorm fine-tuning,
        saliency) therefore builds the graph through the decomposition, as the reference does."""
        for t in tensors:
            if t is not None and not t.is_cuda:
                raise ops.PttError("%s runs on CUDA only (there is no CPU path)" % type(self).__name__)
        if self.training:
            return False
        if not torch.is_grad_enabled():
            return True
        if any(t is not None and t.requires_grad for t in tensors):
            return False
        return not any(p.requires_grad for p in self.parameters())


class PointnetSAModuleVotes(_PackCache, nn.Module):
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
        self.native_train = True      # train() mode runs the native kernels; False: the reference's decomposition in torch

    def _bn_everywhere(self):
        # what the native training kernels cover: conv (no bias) + BatchNorm + ReLU layers, widths in float4 granules
        return all(hasattr(u, "normlayer") and u.conv.bias is None and u.conv.out_channels % 4 == 0 and
                   u.conv.out_channels <= 1024 for u in self.mlp_module)

    def _build_packed(self):
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
        return ops.PackedSAMlp(ws, scales, shifts)

    def _pack(self):
        return self._cached(self._build_packed)

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
        if self.use_xyz and self._use_fused(xyz, features):
            new_xyz = torch.gather(xyz, 1, inds32.long().unsqueeze(-1).expand(-1, -1, 3))
            idx = ops.ball_query(new_xyz, xyz, self.radius, self.nsample)
            feats_pm = ops.cm_to_pm(features.contiguous()) if features is not None else None
            _, new_features = ops.sa_mlp_fwd(self._pack(), xyz, feats_pm, new_xyz, idx, self.radius, self.normalize_xyz,
                                             want_pm=False, want_cm=True)
            return new_xyz, new_features, inds.to(torch.int64)

        xyz_flipped = xyz.transpose(1, 2).contiguous()
        new_xyz = _Gather.apply(xyz_flipped, inds32).transpose(1, 2).contiguous()
        idx = ops.ball_query(new_xyz.detach(), xyz.detach(), self.radius, self.nsample)
        if self.native_train and self.use_xyz and self.training and self._bn_everywhere():
            # native training path (ptt_b200/train_ops.py): tcgen05 contractions, two-phase BatchNorm statistics, fused
            # BatchNorm / ReLU / max-pool backward -- forward AND backward are libptt_b200 kernels
            from . import train_ops
            new_features = train_ops.sa_train(xyz, features, new_xyz, idx, self.radius, self.normalize_xyz, self.mlp_module)
            return new_xyz, new_features, inds.to(torch.int64)
        # the reference's decomposition (pointnet2_modules.py:62-90, pointnet2_utils.py:320-380) under autograd
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


class TransformerBlock(_PackCache, nn.Module):
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
        self.native_train = True      # train() mode runs the native kernels; False: the reference's decomposition in torch

    def _pack(self):
        return self._cached(lambda: ops.PackedTransformer({k: v.detach() for k, v in self.state_dict().items()},
                                                          self.k, self.VARIANT))

    def forward(self, xyz, features):
        if not xyz.is_cuda:
            raise ops.PttError("ptt_b200.TransformerBlock runs on CUDA only (there is no CPU path)")
        xyz = xyz.contiguous()
        features = features.contiguous()
        if self._use_fused(xyz, features):
            r = ops.transformer_block_fwd(self._pack(), xyz, features, want_attn=self.return_attn)
            return r if self.return_attn else (r, None)

        if self.training and self.native_train and self.VARIANT == 0:
            from . import train_ops
            if train_ops.transformer_train_supported(self, xyz, features):
                return train_ops.transformer_train(self, xyz, features)      # forward AND backward on libptt_b200 kernels
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


class TransformerBlockSTD(_PackCache, nn.Module):
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

    def forward(self, xyz, features):
        if not xyz.is_cuda:
            raise ops.PttError("ptt_b200.TransformerBlockSTD runs on CUDA only (there is no CPU path)")
        xyz, features = xyz.contiguous(), features.contiguous()
        if self._use_fused(xyz, features):
            packed = self._cached(lambda: ops.PackedTransformerSTD({k: v.detach() for k, v in self.state_dict().items()}))
            return ops.transformer_std_fwd(packed, xyz, features, want_attn=True)
        pre = features
        x = self.fc1(features)
        q, k, v = self.w_qs(x), self.w_ks(x), self.w_vs(x)
        attn = F.softmax(q @ k.transpose(1, 2) / math.sqrt(k.size(-1)), dim=-1)
        res = attn @ (v + self.fc_delta(xyz))
        return self.fc2(res) + pre, attn


# ----------------------------------------------------------------------------------------------------------------------
# The other blocks of transformer_block.__all__.  Same constructors, forward signatures
# and state_dict keys as the reference classes; in eval / no-grad mode every one of them runs on the kNN vector-attention
# core (ptt_transformer_block_fwd_ex) or on the token-level kernels, otherwise on the reference's decomposition in torch
# over our kNN kernel (autograd).
# ----------------------------------------------------------------------------------------------------------------------
def _take(points, knn_idx):
    B, n, k = knn_idx.shape
    flat = knn_idx.reshape(B, n * k, 1).expand(-1, -1, points.shape[-1])
    return torch.gather(points, 1, flat).reshape(B, n, k, -1)


def _vector_attention(q, kk, v, pos, fc_gamma):
    attn = fc_gamma(q[:, :, None] - kk + pos)
    attn = F.softmax(attn / math.sqrt(kk.size(-1)), dim=-2)
    return torch.einsum("bmnf,bmnf->bmf", attn, v + pos), attn


class _PackedModule(_PackCache, nn.Module):
    def _sd(self):
        return {k: v.detach() for k, v in self.state_dict().items()}

    def _fused(self, *tensors):
        if not self._use_fused(*tensors):
            return False
        key = self._param_key()            # the blocks below build their images inline under `if self._packed is None`
        if self._packed_key != key:
            self._packed, self._packed_key = None, key
        return True

    def _core_layers(self, d_points, d_model, gamma_dim=None):
        g = gamma_dim or d_model
        self.fc_delta = nn.Sequential(nn.Linear(3, d_model), nn.ReLU(), nn.Linear(d_model, d_model))
        self.fc_gamma = nn.Sequential(nn.Linear(g, g), nn.ReLU(), nn.Linear(g, g))
        self.w_qs = nn.Linear(d_model, d_model, bias=False)
        self.w_ks = nn.Linear(d_model, d_model, bias=False)
        self.w_vs = nn.Linear(d_model, d_model, bias=False)


def _core_sd(sd, **override):
    out = {k: sd[k] for k in ops.TRANSFORMER_KEYS if k in sd}
    out.update(override)
    return out


class TransformerBlockMLP(_PackedModule):
    """variants.py:211-256: TransformerBlock with two-layer fc1 / fc2."""

    def __init__(self, d_points, d_model, k, **kwargs):
        super().__init__()
        self.fc1 = nn.Sequential(nn.Linear(d_points, d_model), nn.ReLU(), nn.Linear(d_model, d_model))
        self.fc2 = nn.Sequential(nn.Linear(d_model, d_model), nn.ReLU(), nn.Linear(d_model, d_points))
        self._core_layers(d_points, d_model)
        self.k = k

    def forward(self, xyz, features):
        xyz, features = xyz.contiguous(), features.contiguous()
        B, n, dp = features.shape
        if self._fused(xyz, features):
            if self._packed is None:
                sd = self._sd()
                core = ops.PackedTransformer(_core_sd(sd, **{"fc1.weight": sd["fc1.2.weight"], "fc1.bias": sd["fc1.2.bias"],
                                                             "fc2.weight": sd["fc2.0.weight"], "fc2.bias": sd["fc2.0.bias"]}), self.k)
                self._packed = (ops.PackedLinear(sd["fc1.0.weight"].contiguous(), sd["fc1.0.bias"].contiguous()), core,
                                ops.PackedLinear(sd["fc2.0.weight"].contiguous(), sd["fc2.0.bias"].contiguous()),
                                ops.PackedLinear(sd["fc2.2.weight"].contiguous(), sd["fc2.2.bias"].contiguous()))
            fc1_0, core, fc2_0, fc2_2 = self._packed
            f2 = features.reshape(B * n, dp)
            h = fc1_0(f2, relu=True).reshape(B, n, -1)
            res, attn = ops.transformer_block_fwd_ex(core, xyz, h, flags=2, want_attn=True)
            out = fc2_2(fc2_0(res.reshape(B * n, -1), relu=True), residual=f2).reshape(B, n, dp)
            return out, attn
        knn_idx = ops.knn(xyz.detach(), self.k).long()
        x = self.fc1(features)
        res, attn = _vector_attention(self.w_qs(x), _take(self.w_ks(x), knn_idx), _take(self.w_vs(x), knn_idx),
                                      self.fc_delta(xyz[:, :, None] - _take(xyz, knn_idx)), self.fc_gamma)
        return self.fc2(res) + features, attn


class TransformerBlockCosine(_PackedModule):
    """variants.py:43-88: the attention input is fc_sim([cos(q_i, k_j) | q_i - k_j]) + pos.  fc_sim and fc_gamma.0 are
    linear, so the block is the plain core with w_qs / w_ks pre-multiplied by fc_sim's difference columns plus a rank-1
    term cos(q_i, k_j) * (Wg0 . w_sim) in fc_gamma.0's pre-activation (d_model in {64,128,256,512})."""

    def __init__(self, d_points, d_model, k, **kwargs):
        super().__init__()
        self.fc1 = nn.Linear(d_points, d_model)
        self.fc2 = nn.Linear(d_model, d_points)
        self._core_layers(d_points, d_model)
        self.k = k
        self.fc_sim = nn.Linear(d_model + 1, d_model)

    def forward(self, xyz, features):
        xyz, features = xyz.contiguous(), features.contiguous()
        B, n, dp = features.shape
        if self._fused(xyz, features):
            if self._packed is None:
                sd = {k: v.double() for k, v in self._sd().items()}
                ws, bs = sd["fc_sim.weight"][:, 1:], sd["fc_sim.bias"]
                wg0 = sd["fc_gamma.0.weight"]
                f32 = lambda t: t.float().contiguous()
                core = ops.PackedTransformer(_core_sd(self._sd(), **{
                    "w_qs.weight": f32(ws @ sd["w_qs.weight"]), "w_ks.weight": f32(ws @ sd["w_ks.weight"]),
                    "fc_gamma.0.bias": f32(sd["fc_gamma.0.bias"] + wg0 @ bs)}), self.k)
                lin_q = ops.PackedLinear(f32(sd["w_qs.weight"] @ sd["fc1.weight"]), f32(sd["w_qs.weight"] @ sd["fc1.bias"]))
                lin_k = ops.PackedLinear(f32(sd["w_ks.weight"] @ sd["fc1.weight"]), f32(sd["w_ks.weight"] @ sd["fc1.bias"]))
                self._packed = (core, lin_q, lin_k, f32(wg0 @ sd["fc_sim.weight"][:, 0]))
            core, lin_q, lin_k, vec = self._packed
            knn_idx = ops.knn(xyz, self.k)
            f2 = features.reshape(B * n, dp)
            sim = ops.pair_cosine(lin_q(f2).reshape(B, n, -1), lin_k(f2).reshape(B, n, -1), knn_idx)
            return ops.transformer_block_fwd_ex(core, xyz, features, knn_idx=knn_idx, pair_scalar=sim, pair_vec=vec,
                                                want_attn=True)
        knn_idx = ops.knn(xyz.detach(), self.k).long()
        x = self.fc1(features)
        q, kk, v = self.w_qs(x), _take(self.w_ks(x), knn_idx), _take(self.w_vs(x), knn_idx)
        pos = self.fc_delta(xyz[:, :, None] - _take(xyz, knn_idx))
        sim = F.cosine_similarity(q.unsqueeze(-2).repeat(1, 1, self.k, 1), kk, dim=-1)
        rel = self.fc_sim(torch.cat((sim.unsqueeze(-1), q[:, :, None] - kk), dim=-1))
        attn = F.softmax(self.fc_gamma(rel + pos) / math.sqrt(kk.size(-1)), dim=-2)
        res = torch.einsum("bmnf,bmnf->bmf", attn, v + pos)
        return self.fc2(res) + features, attn


class TransformerBlockALL(_PackedModule):
    """variants.py:91-124: no neighbourhoods -- per-token gate, softmax over ALL n tokens of a cloud per channel."""

    def __init__(self, d_points, d_model, k, **kwargs):
        super().__init__()
        self.fc1 = nn.Linear(d_points, d_model)
        self.fc2 = nn.Linear(d_model, d_points)
        self._core_layers(d_points, d_model)
        self.k = k

    def forward(self, xyz, features):
        xyz, features = xyz.contiguous(), features.contiguous()
        B, n, dp = features.shape
        if self._fused(xyz, features):
            if self._packed is None:
                sd = {k: v.double() for k, v in self._sd().items()}
                f32 = lambda t: t.float().contiguous()
                lin = lambda w, b=None: ops.PackedLinear(f32(w), f32(b) if b is not None else None)
                wd = sd["w_qs.weight"] - sd["w_ks.weight"]
                self._packed = dict(
                    qk=lin(wd @ sd["fc1.weight"], wd @ sd["fc1.bias"]),                       # (Wq - Wk) fc1
                    v=lin(sd["w_vs.weight"] @ sd["fc1.weight"], sd["w_vs.weight"] @ sd["fc1.bias"]),
                    d0=lin(sd["fc_delta.0.weight"], sd["fc_delta.0.bias"]), d2=lin(sd["fc_delta.2.weight"], sd["fc_delta.2.bias"]),
                    g0=lin(sd["fc_gamma.0.weight"], sd["fc_gamma.0.bias"]), g2=lin(sd["fc_gamma.2.weight"], sd["fc_gamma.2.bias"]),
                    fc2=lin(sd["fc2.weight"], sd["fc2.bias"]))
            P = self._packed
            f2 = features.reshape(B * n, dp)
            pos = P["d2"](P["d0"](xyz.reshape(B * n, 3), relu=True))
            logits = P["g2"](P["g0"](P["qk"](f2, residual=pos), relu=True))
            dm = logits.shape[1]
            res, attn = ops.token_softmax_gate(logits.reshape(B, n, dm), P["v"](f2, residual=pos).reshape(B, n, dm),
                                               math.sqrt(dm), want_attn=True)
            return P["fc2"](res.reshape(B * n, dm), residual=f2).reshape(B, n, dp), attn
        x = self.fc1(features)
        q, kk, v = self.w_qs(x), self.w_ks(x), self.w_vs(x)
        pos = self.fc_delta(xyz)
        attn = F.softmax(self.fc_gamma(q - kk + pos) / math.sqrt(kk.size(-1)), dim=-2)
        return self.fc2(attn * (v + pos)) + features, attn


class CrossAttentionBlock(_PackedModule):
    """variants.py:168-208: queries from the template features, keys / values from the search features (fc2 is defined
    but unused by the reference; fc3 is the output projection)."""

    def __init__(self, d_points, d_model, k, **kwargs):
        super().__init__()
        self.fc1 = nn.Linear(d_points, d_model)
        self.fc2 = nn.Linear(d_points, d_model)
        self.fc3 = nn.Linear(d_model, d_points)
        self._core_layers(d_points, d_model)
        self.k = k

    def forward(self, xyz, search_feat, template_feat):
        xyz, search_feat, template_feat = xyz.contiguous(), search_feat.contiguous(), template_feat.contiguous()
        if self._fused(xyz, search_feat, template_feat):
            if self._packed is None:
                sd = self._sd()
                self._packed = ops.PackedTransformer(_core_sd(sd, **{"fc2.weight": sd["fc3.weight"], "fc2.bias": sd["fc3.bias"]}), self.k)
            return ops.transformer_block_fwd_ex(self._packed, xyz, search_feat, q_features=template_feat, want_attn=True)
        knn_idx = ops.knn(xyz.detach(), self.k).long()
        s, tm = self.fc1(search_feat), self.fc1(template_feat)
        res, attn = _vector_attention(self.w_qs(tm), _take(self.w_ks(s), knn_idx), _take(self.w_vs(s), knn_idx),
                                      self.fc_delta(xyz[:, :, None] - _take(xyz, knn_idx)), self.fc_gamma)
        return self.fc3(res) + search_feat, attn


class TransformerBlockBackbone(_PackedModule):
    """variants.py:259-294: the attention core over EXTERNALLY supplied neighbourhoods (ball-query groups), returning the
    aggregated d_model features (no fc2, no residual; the reference's debug prints are not reproduced).  The fused path
    covers the case the arithmetic of the reference admits -- one query per point (features.shape[1] == npoint) with
    grouped_xyz gathered from new_xyz by grouped_idx; anything else runs on the decomposition."""

    def __init__(self, d_points, d_model, k, **kwargs):
        super().__init__()
        self.fc1 = nn.Linear(d_points, d_model)
        self.fc2 = nn.Linear(d_model, d_points)
        self._core_layers(d_points, d_model)
        self.k = k

    def forward(self, new_xyz, grouped_xyz, grouped_idx, features):
        new_xyz, features = new_xyz.contiguous(), features.contiguous()
        gx = grouped_xyz.permute(0, 2, 3, 1).contiguous()
        idx = grouped_idx.long()
        if self._fused(new_xyz, features) and features.shape[1] == new_xyz.shape[1] and torch.equal(_take(new_xyz, idx), gx):
            if self._packed is None:
                self._packed = ops.PackedTransformer(_core_sd(self._sd()), grouped_idx.shape[2])
            self._packed.k = grouped_idx.shape[2]
            return ops.transformer_block_fwd_ex(self._packed, new_xyz, features, flags=2, knn_idx=grouped_idx.int().contiguous())
        x = self.fc1(features)
        res, _ = _vector_attention(self.w_qs(x), _take(self.w_ks(x), idx), _take(self.w_vs(x), idx),
                                   self.fc_delta(new_xyz[:, :, None] - gx), self.fc_gamma)
        return res.contiguous()


class MulHeadTransformerLayer(_PackedModule):
    """multitransformer.py:11-63.  fc_gamma acts on head_dim slices with weights shared by the heads, i.e. it is the
    block-diagonal d_model x d_model map diag(Wg, ..., Wg): the layer is the plain core with those weights and the
    temperature sqrt(head_dim), followed by proj -> LayerNorm -> fc2 -> LayerNorm -> + input."""

    def __init__(self, d_points, d_model, k, heads, drop=0.0):
        super().__init__()
        self.heads = heads
        head_dim = d_model // heads
        self.fc1 = nn.Linear(d_points, d_model)
        self.fc2 = nn.Linear(d_model, d_points)
        self._core_layers(d_points, d_model, gamma_dim=head_dim)
        self.proj = nn.Linear(d_model, d_model, bias=False)
        self.proj_drop = nn.Dropout(drop)
        self.k = k
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_points)

    def forward(self, xyz, features):
        xyz, features = xyz.contiguous(), features.contiguous()
        B, n, dp = features.shape
        H = self.heads
        if self._fused(xyz, features):
            if self._packed is None:
                sd = self._sd()
                bd = lambda w: torch.block_diag(*([w] * H)).contiguous()
                core = ops.PackedTransformer(_core_sd(sd, **{
                    "fc_gamma.0.weight": bd(sd["fc_gamma.0.weight"]), "fc_gamma.0.bias": sd["fc_gamma.0.bias"].repeat(H).contiguous(),
                    "fc_gamma.2.weight": bd(sd["fc_gamma.2.weight"]), "fc_gamma.2.bias": sd["fc_gamma.2.bias"].repeat(H).contiguous()}),
                    self.k)
                self._packed = (core, ops.PackedLinear(sd["proj.weight"].contiguous()),
                                ops.PackedLinear(sd["fc2.weight"].contiguous(), sd["fc2.bias"].contiguous()))
            core, proj, fc2 = self._packed
            dm = core.d_model
            res, attn = ops.transformer_block_fwd_ex(core, xyz, features, flags=2, divisor=math.sqrt(dm // H), want_attn=True)
            y = ops.layer_norm(proj(res.reshape(B * n, dm)), self.norm1.weight.detach(), self.norm1.bias.detach(), self.norm1.eps)
            y = ops.layer_norm(fc2(y), self.norm2.weight.detach(), self.norm2.bias.detach(), self.norm2.eps,
                               residual=features.reshape(B * n, dp))
            attn = attn.view(B, n, self.k, H, dm // H).permute(0, 3, 1, 2, 4).flatten(0, 1)     # the reference's (B*H, n, k, head_dim)
            return y.reshape(B, n, dp), attn
        knn_idx = ops.knn(xyz.detach(), self.k).long()
        x = self.fc1(features)
        C = x.shape[2]
        query = self.w_qs(x).view(B, n, H, -1).permute(0, 2, 1, 3).flatten(0, 1)
        split = lambda t: t.view(B, n, t.shape[2], H, -1).permute(0, 3, 1, 2, 4).flatten(0, 1)
        pos, key, value = map(split, (self.fc_delta(xyz[:, :, None] - _take(xyz, knn_idx)), _take(self.w_ks(x), knn_idx),
                                      _take(self.w_vs(x), knn_idx)))
        res, attn = _vector_attention(query, key, value, pos, self.fc_gamma)
        if H > 1:
            res = res.permute(0, 2, 1).reshape(B, C, n).permute(0, 2, 1)
        res = self.norm1(self.proj_drop(self.proj(res)))
        return self.norm2(self.fc2(res)) + features, attn


class MulTransformerBlock(nn.Module):
    """multitransformer.py:66-76: `layers` deep copies of one MulHeadTransformerLayer applied in sequence."""

    def __init__(self, d_points, d_model, k, heads, layers):
        super().__init__()
        import copy
        layer = MulHeadTransformerLayer(d_points, d_model, k, heads)
        self.layers = nn.ModuleList([copy.deepcopy(layer) for _ in range(layers)])

    def forward(self, xyz, features):
        output, attn = features, None
        for layer in self.layers:
            output, attn = layer(xyz, output)
        return output, attn


REGISTRY = {
    "MulTransformerBlock": MulTransformerBlock, "TransformerBlock": TransformerBlock, "TransformerBlockALL": TransformerBlockALL,
    "TransformerBlockBackbone": TransformerBlockBackbone, "TransformerBlockCosine": TransformerBlockCosine,
    "TransformerBlockMLP": TransformerBlockMLP, "TransformerBlockOffset": TransformerBlockOffset,
    "TransformerBlockSTD": TransformerBlockSTD, "CrossAttentionBlock": CrossAttentionBlock,
}


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
    for name, cls in REGISTRY.items():          # every name of transformer_block/__init__.py:7-17
        transformer_block.__all__[name] = cls
    return pointnet2_modules, transformer_block
