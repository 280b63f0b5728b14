"""Native training path of the hot path's modules: autograd.Functions whose forward AND backward are libptt_b200
kernels (BASELINE configs[3]; VERDICT r1 "missing" #1).

`sa_train`   one set-abstraction layer in train() mode -- QueryAndGroup -> [Conv2d 1x1 -> BatchNorm2d(batch statistics)
             -> ReLU] x L -> max over nsample (pointnet2_modules.py:57-90, pytorch_utils.py:12-36) -- with its backward.

Design (csrc/ws_gemm.cu, train_ops.cu, tc_gemm.cu, tc_wgrad.cu): activations are pair-row matrices; each layer is one tcgen05
contraction that writes the PRE-BatchNorm output y_l once and leaves the column statistics of y_l in its epilogue; the normalised,
rectified activation is never stored -- the next contraction (and the weight-gradient contraction in the backward pass)
apply relu(ka * y + kb) while loading y_l.  Saved for backward: the grouped input rows, y_l per layer, the per-channel
(ka, kb, mean, rstd) vectors and the max-pool indices.
"""
import ctypes

import numpy as np
import torch

from . import _lib, ops
from ._lib import check
from .ops import PttError, _DeviceGuard, _F, _I, _ptr, _req, _stream

_F64 = torch.float64


def _pad4(c):
    return (int(c) + 3) // 4 * 4


def linear_wgrad(dy, x, M, N, x_affine=None, want_bias=False):
    """dW (M, N) = dy[:, :M]^T . f(x[:, :N]), f = identity or relu(ka * x + kb) (ptt_linear_wgrad).
    want_bias: -> (dW, db), db (M,) = the column sums of dy, accumulated by the same kernel."""
    _req(dy, _F, 2, "dy"), _req(x, _F, 2, "x")
    R = dy.shape[0]
    if x.shape[0] != R or dy.shape[1] < M or x.shape[1] < N:
        raise PttError("linear_wgrad: inconsistent shapes")
    ka, kb = x_affine if x_affine is not None else (None, None)
    ldw = _pad4(N)
    with _DeviceGuard(dy.device):
        buf = torch.zeros(M * ldw + (M if want_bias else 0), dtype=_F, device=dy.device)   # one fill for both results
        dw = buf[:M * ldw].view(M, ldw)
        db = buf[M * ldw:] if want_bias else None
        check(_lib.lib().ptt_linear_wgrad(_ptr(dy), dy.shape[1], _ptr(x), x.shape[1], _ptr(ka), _ptr(kb), R, int(M), int(N),
                                          _ptr(dw), ldw, _ptr(db), _stream()), "ptt_linear_wgrad")
    return (dw[:, :N], db) if want_bias else dw[:, :N]


def col_stats(y, C):
    _req(y, _F, 2, "y")
    with _DeviceGuard(y.device):
        sums = torch.empty(2, C, dtype=_F64, device=y.device)
        check(_lib.lib().ptt_col_stats(_ptr(y), y.shape[1], y.shape[0], int(C), _ptr(sums), _stream()), "ptt_col_stats")
    return sums


def bn_train_finalize(sums, R, gamma, beta, eps, momentum, running_mean, running_var):
    """-> (ka, kb, mean, rstd); running statistics updated in place."""
    C = sums.shape[1]
    with _DeviceGuard(sums.device):
        vec = torch.empty(4, C, dtype=_F, device=sums.device)
        check(_lib.lib().ptt_bn_train_finalize(_ptr(sums), int(R), C, _ptr(gamma), _ptr(beta), float(eps), float(momentum),
                                               _ptr(running_mean), _ptr(running_var), _ptr(vec[0]), _ptr(vec[1]), _ptr(vec[2]),
                                               _ptr(vec[3]), _stream()), "ptt_bn_train_finalize")
    return vec[0], vec[1], vec[2], vec[3]


def sa_group_rows(xyz, feats_pm, new_xyz, idx, C, radius, normalize):
    B, N, _ = xyz.shape
    _, M, ns = idx.shape
    ld = _pad4(C + 3)
    with _DeviceGuard(xyz.device):
        rows = torch.empty(B * M * ns, ld, dtype=_F, device=xyz.device)
        check(_lib.lib().ptt_sa_group_rows(_ptr(xyz), _ptr(feats_pm), feats_pm.shape[2] if feats_pm is not None else 0,
                                           _ptr(new_xyz), _ptr(idx), B, N, M, ns, int(C), float(radius), int(bool(normalize)),
                                           _ptr(rows), ld, _stream()), "ptt_sa_group_rows")
    return rows


def sa_group_rows_grad(d_rows, idx, N, C, radius, normalize, want_feats, want_xyz):
    B, M, ns = idx.shape
    with _DeviceGuard(d_rows.device):
        d_feats = torch.zeros(B, N, _pad4(C), dtype=_F, device=d_rows.device) if want_feats and C > 0 else None
        d_xyz = torch.zeros(B, N, 3, dtype=_F, device=d_rows.device) if want_xyz else None
        d_new = torch.zeros(B, M, 3, dtype=_F, device=d_rows.device) if want_xyz else None
        check(_lib.lib().ptt_sa_group_rows_grad(_ptr(d_rows), d_rows.shape[1], _ptr(idx), B, int(N), M, ns, int(C), float(radius),
                                                int(bool(normalize)), _ptr(d_feats), d_feats.shape[2] if d_feats is not None else 0,
                                                _ptr(d_xyz), _ptr(d_new), _stream()), "ptt_sa_group_rows_grad")
    return d_feats, d_xyz, d_new


def bn_relu_maxpool(y, groups, ns, C, ka, kb):
    with _DeviceGuard(y.device):
        out = torch.empty(groups, C, dtype=_F, device=y.device)
        arg = torch.empty(groups, C, dtype=_I, device=y.device)
        check(_lib.lib().ptt_bn_relu_maxpool(_ptr(y), y.shape[1], int(groups), int(ns), int(C), _ptr(ka), _ptr(kb), _ptr(out), C,
                                             _ptr(arg), _stream()), "ptt_bn_relu_maxpool")
    return out, arg


def bn_relu_bwd(dz, argmax, ns, y, C, ka, kb, mean, rstd, gamma):
    """-> (dy (R, C), sums (2, C) float64 = (d beta, d gamma), dparam (2, C) float32 = the same two rows)."""
    R = y.shape[0]
    with _DeviceGuard(y.device):
        dy = torch.empty(R, C, dtype=_F, device=y.device)
        sums = torch.empty(2, C, dtype=_F64, device=y.device)
        dparam = torch.empty(2, C, dtype=_F, device=y.device)
        check(_lib.lib().ptt_bn_relu_bwd(_ptr(dz), dz.shape[1], _ptr(argmax), int(ns), _ptr(y), y.shape[1], R, int(C), _ptr(ka),
                                         _ptr(kb), _ptr(mean), _ptr(rstd), _ptr(gamma), _ptr(sums), _ptr(dy), C, _ptr(dparam),
                                         _stream()), "ptt_bn_relu_bwd")
    return dy, sums, dparam


class TrainPacks:
    """Per-module cache of the packed weight images the native training path consumes (the layer's weight for the forward
    contraction, its transpose for the input gradient).  An image is valid while the parameter's `_version` and address
    are the ones it was packed from: `get` repacks a stale image with one launch, `repack_stale` refreshes the stale images
    of MANY modules with ONE launch (ptt_linear_pack_batch) -- a training step calls it once after the optimiser update."""

    def __init__(self):
        self.entries = {}

    @staticmethod
    def _ver(weight, bias):
        return (weight._version, weight.data_ptr(), -1 if bias is None else bias._version, 0 if bias is None else bias.data_ptr())

    @staticmethod
    def _src(weight, shape2d, transposed):
        w2 = weight.detach().reshape(shape2d)
        return w2.t() if transposed else w2

    def get(self, key, weight, shape2d, transposed=False, bias=None):
        ver = self._ver(weight, bias)
        ent = self.entries.get(key)
        if ent is None or ent["shape"] != tuple(shape2d):
            packed = ops.PackedLinear(self._src(weight, shape2d, transposed), None if bias is None else bias.detach(), check_range=False)
            ent = {"packed": packed, "shape": tuple(shape2d), "t": bool(transposed)}
            self.entries[key] = ent
        elif ent["ver"] != ver:
            ent["packed"].repack(self._src(weight, shape2d, transposed), None if bias is None else bias.detach())
        ent["ver"], ent["weight"], ent["bias"] = ver, weight, bias
        return ent["packed"]


def invalidate_packs(modules):
    """Mark every cached image of `modules` stale.  Needed after CUDA-graph REPLAYS of a training step: a replay updates the
    parameters on the device without touching their Python-side `_version`, so the caches cannot see that the images of the
    last replay were packed from the parameters BEFORE its optimiser update."""
    for m in modules:
        packs = getattr(m, "train_packs", None)
        if packs is not None:
            for ent in packs.entries.values():
                ent["ver"] = None


_DESC = np.dtype([("w", "<u8"), ("ld_c", "<i8"), ("ld_k", "<i8"), ("b", "<u8"), ("p", "<u8"), ("K", "<i4"), ("C", "<i4")])   # PttPackDesc
_TABLES = {}          # (device, descriptor bytes) -> device copy of the table (a step's set of stale layers repeats every step)


def repack_stale(modules):
    """Refresh, with one launch, every stale image of the `train_packs` caches of `modules` (no-op while fewer than two are
    stale; the per-image path of TrainPacks.get handles whatever this leaves)."""
    stale = []
    for m in modules:
        packs = getattr(m, "train_packs", None)
        if packs is None:
            continue
        for ent in packs.entries.values():
            if ent["ver"] != TrainPacks._ver(ent["weight"], ent["bias"]):
                stale.append(ent)
    if len(stale) < 2:
        return 0
    dev = stale[0]["weight"].device
    desc = np.zeros(len(stale), dtype=_DESC)
    for i, ent in enumerate(stale):
        w, cout, k = ent["weight"], ent["packed"].cout, ent["packed"].k
        rows, cols = ent["shape"]                              # the (rows, cols) matrix the parameter is viewed as
        desc[i] = ((w.data_ptr(), 1, cols, 0, 0, k, cout) if ent["t"] else (w.data_ptr(), cols, 1, 0, 0, k, cout))
        desc[i]["b"] = 0 if ent["bias"] is None else ent["bias"].data_ptr()
        desc[i]["p"] = ent["packed"].params.data_ptr()
    key = (str(dev), desc.tobytes())
    table = _TABLES.get(key)
    if table is None:
        if torch.cuda.is_current_stream_capturing():
            return 0                                           # no host -> device copy inside a capture; the per-image path runs
        if len(_TABLES) > 64:
            _TABLES.clear()
        table = torch.from_numpy(desc.view(np.uint8).copy()).to(dev)
        _TABLES[key] = table
    with _DeviceGuard(dev):
        check(_lib.lib().ptt_linear_pack_batch(_ptr(table), len(stale), _stream()), "ptt_linear_pack_batch")
    for ent in stale:
        ent["packed"].has_bias = ent["bias"] is not None
        ent["ver"] = TrainPacks._ver(ent["weight"], ent["bias"])
    return len(stale)


class _SATrain(torch.autograd.Function):
    """inputs: xyz (B,N,3), feats_cm (B,C,N) | None, new_xyz (B,M,3), idx (B,M,ns) int32, then per layer
    (conv weight (Cout,Cin,1,1), bn weight, bn bias, running_mean, running_var)."""

    @staticmethod
    def forward(ctx, xyz, feats_cm, new_xyz, idx, radius, normalize, eps, momentum, packs, *params):
        L = len(params) // 5
        B, N, _ = xyz.shape
        _, M, ns = idx.shape
        C = feats_cm.shape[1] if feats_cm is not None else 0
        feats_pm = ops.cm_to_pm(feats_cm.contiguous(), _pad4(C)) if C else None
        x0 = sa_group_rows(xyz, feats_pm, new_xyz, idx, C, radius, normalize)
        R = x0.shape[0]
        ys, affs, stats = [], [], []
        src, aff, k_in = x0, None, C + 3
        for l in range(L):
            w, g, b, rm, rv = params[5 * l: 5 * l + 5]
            cout = w.shape[0]
            lin = packs.get((l, "fwd"), w, (cout, k_in))
            y, sums = ops.linear_with_stats(lin, src, in_affine=aff)      # the statistics come out of the contraction's epilogue
            ka, kb, mean, rstd = bn_train_finalize(sums, R, g.detach(), b.detach(), eps, momentum, rm, rv)
            ys.append(y)
            affs.append((ka, kb))
            stats.append((mean, rstd))
            src, aff, k_in = y, (ka, kb), cout
        out, arg = bn_relu_maxpool(ys[-1], B * M, ns, k_in, *affs[-1])
        ctx.save_for_backward(x0, idx, arg, *ys, *[t for a in affs for t in a], *[t for s in stats for t in s],
                              *[params[5 * l].detach() for l in range(L)], *[params[5 * l + 1].detach() for l in range(L)])
        ctx.meta = (L, B, N, M, ns, C, float(radius), bool(normalize))
        ctx.packs = packs
        ctx.mark_non_differentiable(arg)
        return ops.pm_to_cm(out.view(B, M, k_in)), arg

    @staticmethod
    def backward(ctx, grad_out, _grad_arg):
        L, B, N, M, ns, C, radius, normalize = ctx.meta
        sv = ctx.saved_tensors
        x0, idx, arg = sv[0], sv[1], sv[2]
        ys = sv[3: 3 + L]
        affs = [(sv[3 + L + 2 * l], sv[3 + L + 2 * l + 1]) for l in range(L)]
        stats = [(sv[3 + 3 * L + 2 * l], sv[3 + 3 * L + 2 * l + 1]) for l in range(L)]
        ws = sv[3 + 5 * L: 3 + 6 * L]
        gs = sv[3 + 6 * L: 3 + 7 * L]
        grads = [None] * (5 * L)
        dz = ops.cm_to_pm(grad_out.contiguous())                     # (B, M, C_L) -> pooled rows (B*M, C_L)
        dz = dz.view(B * M, dz.shape[2])
        pooled = arg
        for l in range(L - 1, -1, -1):
            cout, cin = ws[l].shape[0], ws[l].shape[1]
            dy, _, dparam = bn_relu_bwd(dz, pooled, ns, ys[l], cout, affs[l][0], affs[l][1], stats[l][0], stats[l][1], gs[l])
            pooled = None
            grads[5 * l + 1] = dparam[1]                             # d gamma
            grads[5 * l + 2] = dparam[0]                             # d beta
            src, aff = (x0, None) if l == 0 else (ys[l - 1], affs[l - 1])
            grads[5 * l] = linear_wgrad(dy, src, cout, cin, aff).reshape(ws[l].shape)
            need_dx = l > 0 or ctx.needs_input_grad[0] or (C > 0 and ctx.needs_input_grad[1]) or ctx.needs_input_grad[2]
            if need_dx:
                wt = ctx.packs.get((l, "dgrad"), ws[l], (cout, cin), transposed=True)   # strided pack: no transposed copy
                # padded rows only below layer 0, where ptt_sa_group_rows_grad reads the first 3 + C columns and nothing else
                dz, _ = ops.linear_with_stats(wt, dy, want_stats=False, ld_out=_pad4(cin), zero_pad=False)
        d_xyz = d_feats = d_new = None
        if ctx.needs_input_grad[0] or (C > 0 and ctx.needs_input_grad[1]) or ctx.needs_input_grad[2]:
            want_xyz = ctx.needs_input_grad[0] or ctx.needs_input_grad[2]
            d_feats_pm, d_xyz, d_new = sa_group_rows_grad(dz, idx, N, C, radius, normalize, C > 0 and ctx.needs_input_grad[1], want_xyz)
            if d_feats_pm is not None:
                d_feats = ops.pm_to_cm(d_feats_pm, C)
        return (d_xyz, d_feats, d_new, None, None, None, None, None, None, *grads)


def sa_train(xyz, feats_cm, new_xyz, idx, radius, normalize, mlp_module):
    """The training-mode body of PointnetSAModuleVotes on the native path.  mlp_module: the reference-shaped SharedMLP
    (layers `layer{i}.conv` / `.normlayer.bn`); its BatchNorm running statistics are updated in place."""
    params, eps, momentum = [], None, None
    for unit in mlp_module:
        if not hasattr(unit, "normlayer") or unit.conv.bias is not None:
            raise PttError("sa_train: conv (no bias) + BatchNorm layers only (pytorch_utils.py:25-36 with bn=True)")
        bn = unit.normlayer.bn
        if eps is None:
            eps, momentum = bn.eps, (bn.momentum if bn.momentum is not None else 0.1)
        elif (bn.eps, bn.momentum if bn.momentum is not None else 0.1) != (eps, momentum):
            raise PttError("sa_train: the BatchNorm layers of one SharedMLP share eps / momentum")
        params += [unit.conv.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var]
        bn.num_batches_tracked += 1
    if getattr(mlp_module, "train_packs", None) is None:
        mlp_module.train_packs = TrainPacks()
    out, _ = _SATrain.apply(xyz, feats_cm, new_xyz, idx, float(radius), bool(normalize), float(eps), float(momentum),
                            mlp_module.train_packs, *params)
    return out


# ------------------------------------------------------------------------------------------------
# kNN vector-attention block (transformer_block/variants.py:127-165) in train() mode
# ------------------------------------------------------------------------------------------------

def _tr_layout(B, n, k, dp, dm):
    out = (ctypes.c_size_t * 7)()
    check(_lib.lib().ptt_transformer_block_workspace_layout(B, n, k, dp, dm, out), "ptt_transformer_block_workspace_layout")
    return [int(v) for v in out]


def rows_linear_masked(packed, x, mask_ref=None):
    """out (R, d) = x . W^T [* (mask_ref > 0)] for a packed bias-free square layer (ptt_tr_rows_linear: the persistent
    CTA-pair kernel of the forward passes, ReLU backward fused into its epilogue); other shapes fall back to the generic
    contraction + ptt_tr_mask_positive."""
    R, d = x.shape[0], packed.cout
    if packed.k == d and d in (256, 512) and not packed.has_bias and x.shape[1] % 4 == 0:
        with _DeviceGuard(x.device):
            out = torch.empty(R, d, dtype=_F, device=x.device)
            rc = _lib.lib().ptt_tr_rows_linear(_ptr(x), x.shape[1], R, d, _ptr(packed.params), _ptr(mask_ref), _ptr(out), _stream())
        if rc == 0:
            return out
        if rc != _lib.PTT_ERR_UNSUPPORTED:
            check(rc, "ptt_tr_rows_linear")
    out = packed(x)
    if mask_ref is not None:
        with _DeviceGuard(x.device):
            check(_lib.lib().ptt_tr_mask_positive(_ptr(out), _ptr(mask_ref), out.numel(), _stream()), "ptt_tr_mask_positive")
    return out


class _TransformerTrain(torch.autograd.Function):
    """inputs: xyz (B,n,3), features (B,n,d_points), k, then the 15 parameters in ops.TRANSFORMER_KEYS order.
    Forward = the fused eval kernels (there is no BatchNorm in the block) with attn and a kept workspace; backward
    recomputes the cheap token-level projections, reads g = relu(fc_gamma.0(.)) and pos + v from that workspace and runs
    the input / weight gradient contractions on tcgen05 (ptt_linear_fwd, ptt_linear_wgrad)."""

    @staticmethod
    def forward(ctx, xyz, features, k, packs, *params):
        sd = {key: p.detach() for key, p in zip(ops.TRANSFORMER_KEYS, params)}
        packed = ops.PackedTransformer(sd, k, check_range=False)
        B, n, dp = features.shape
        knn = ops.knn(xyz, k)
        with _DeviceGuard(xyz.device):
            ws = torch.empty(packed.workspace_bytes(B, n) // 4 + 16, dtype=_F, device=xyz.device)
        out, attn = ops.transformer_block_fwd(packed, xyz, features, knn_idx=knn, want_attn=True, workspace=ws)
        ctx.save_for_backward(xyz, features, knn, attn, ws, *[p.detach() for p in params])
        ctx.meta = (B, n, int(k), dp, packed.d_model)
        ctx.packs = packs
        ctx.mark_non_differentiable(attn)
        return out, attn

    @staticmethod
    def backward(ctx, dout, _dattn):
        B, n, k, dp, dm = ctx.meta
        sv = ctx.saved_tensors
        xyz, features, knn, attn, ws = sv[:5]
        W = dict(zip(ops.TRANSFORMER_KEYS, sv[5:]))
        tokens, pairs = B * n, B * n * k
        off = _tr_layout(B, n, k, dp, dm)
        ld = off[6]
        res = ws[off[3]: off[3] + tokens * ld].view(tokens, ld)
        g = ws[off[4]: off[4] + pairs * ld].view(pairs, ld)
        vp = ws[off[5]: off[5] + pairs * ld].view(pairs, ld)
        names = {id(t): key for key, t in W.items()}
        lin = lambda w, b=None: ctx.packs.get((names[id(w)], "fwd"), w, tuple(w.shape), bias=b)
        lin_t = lambda w: ctx.packs.get((names[id(w)], "dgrad"), w, tuple(w.shape), transposed=True)   # strided: no transposed copy
        f2 = features.reshape(tokens, dp)
        dout2 = dout.reshape(tokens, dp).contiguous()
        # token-level projections, recomputed (tokens x d_model each)
        x = lin(W["fc1.weight"], W["fc1.bias"])(f2)
        q, kk, v = lin(W["w_qs.weight"])(x), lin(W["w_ks.weight"])(x), lin(W["w_vs.weight"])(x)
        with _DeviceGuard(xyz.device):
            h1 = torch.empty(pairs, ld, dtype=_F, device=xyz.device)
            delta = torch.empty(pairs, 4, dtype=_F, device=xyz.device)
            a_in = torch.empty(pairs, ld, dtype=_F, device=xyz.device)
            L = _lib.lib()
            check(L.ptt_tr_pair_inputs(_ptr(xyz), _ptr(knn), B, n, k, dm, _ptr(W["fc_delta.0.weight"]), _ptr(W["fc_delta.0.bias"]),
                                       _ptr(q), _ptr(kk), _ptr(v), dm, _ptr(vp), ld, _ptr(h1), _ptr(delta), _ptr(a_in), _stream()),
                  "ptt_tr_pair_inputs")
            grads = {}
            # out = fc2(res) + features
            grads["fc2.weight"], grads["fc2.bias"] = linear_wgrad(dout2, res, dp, dm, want_bias=True)
            dres = lin_t(W["fc2.weight"])(dout2)
            dlogit = torch.empty(pairs, ld, dtype=_F, device=xyz.device)
            dvp = torch.empty(pairs, ld, dtype=_F, device=xyz.device)
            check(L.ptt_tr_softmax_bwd(_ptr(dres), dm, _ptr(attn), _ptr(vp), ld, tokens, k, dm, float(dm) ** 0.5, _ptr(dlogit),
                                       _ptr(dvp), _stream()), "ptt_tr_softmax_bwd")
            # logits = fc_gamma.2(g), g = relu(fc_gamma.0(a_in))
            grads["fc_gamma.2.weight"], grads["fc_gamma.2.bias"] = linear_wgrad(dlogit, g, dm, dm, want_bias=True)
            dpre = rows_linear_masked(lin_t(W["fc_gamma.2.weight"]), dlogit, g if ld == dm else None)
            if ld != dm:
                check(L.ptt_tr_mask_positive(_ptr(dpre), _ptr(g), pairs * ld, _stream()), "ptt_tr_mask_positive")
            grads["fc_gamma.0.weight"], grads["fc_gamma.0.bias"] = linear_wgrad(dpre, a_in, dm, dm, want_bias=True)
            da = rows_linear_masked(lin_t(W["fc_gamma.0.weight"]), dpre)
            # a_in = q_i - k_j + pos_ij ; vp = v_j + pos_ij
            dq = torch.empty(tokens, dm, dtype=_F, device=xyz.device)
            dk = torch.zeros(tokens, dm, dtype=_F, device=xyz.device)
            dv = torch.zeros(tokens, dm, dtype=_F, device=xyz.device)
            check(L.ptt_tr_pair_scatter(_ptr(da), _ptr(dvp), ld, _ptr(knn), B, n, k, dm, _ptr(dq), _ptr(dk), _ptr(dv), dm, _stream()),
                  "ptt_tr_pair_scatter")
            dpos = da
            # pos = fc_delta.2(h1), h1 = relu(fc_delta.0(delta))
            grads["fc_delta.2.weight"], grads["fc_delta.2.bias"] = linear_wgrad(dpos, h1, dm, dm, want_bias=True)
            dh1 = rows_linear_masked(lin_t(W["fc_delta.2.weight"]), dpos, h1 if ld == dm else None)
            if ld != dm:
                check(L.ptt_tr_mask_positive(_ptr(dh1), _ptr(h1), pairs * ld, _stream()), "ptt_tr_mask_positive")
            grads["fc_delta.0.weight"], grads["fc_delta.0.bias"] = linear_wgrad(dh1, delta, dm, 3, want_bias=True)
            # q, k, v = W x ; x = fc1(features)
            grads["w_qs.weight"] = linear_wgrad(dq, x, dm, dm)
            grads["w_ks.weight"] = linear_wgrad(dk, x, dm, dm)
            grads["w_vs.weight"] = linear_wgrad(dv, x, dm, dm)
            dx = lin_t(W["w_qs.weight"])(dq)
            dx = lin_t(W["w_ks.weight"])(dk, residual=dx)
            dx = lin_t(W["w_vs.weight"])(dv, residual=dx)
            grads["fc1.weight"], grads["fc1.bias"] = linear_wgrad(dx, f2, dm, dp, want_bias=True)
            df = lin_t(W["fc1.weight"])(dx, residual=dout2)
        return (None, df.view(B, n, dp), None, None, *[grads[key].contiguous() for key in ops.TRANSFORMER_KEYS])


def transformer_train_supported(block, xyz, features):
    dm, dp, k = block.fc1.out_features, block.fc1.in_features, block.k
    return (not xyz.requires_grad and dm in (64, 128, 256, 512) and dp % 4 == 0 and k >= 1 and (k & (k - 1)) == 0 and k <= 32
            and xyz.shape[1] >= k)


def transformer_train(block, xyz, features):
    """TransformerBlock.forward in train() mode on the native path -> (res, attn)."""
    params = [dict(block.named_parameters())[key] for key in ops.TRANSFORMER_KEYS]
    if getattr(block, "train_packs", None) is None:
        block.train_packs = TrainPacks()
    return _TransformerTrain.apply(xyz, features, block.k, block.train_packs, *params)
