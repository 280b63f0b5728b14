"""Tensor-level wrappers of the C ABI (include/ptt_b200.h): argument checks, output allocation, stream.

PyTorch is used for device memory and the current CUDA stream only.  Every function enqueues on
`torch.cuda.current_stream()` and never synchronises.  Inputs must be contiguous CUDA tensors of the
stated dtype (RuntimeError otherwise, like the reference's `_ext`); outputs are fresh tensors.
"""
import ctypes

import torch

from . import _lib
from ._lib import PttError, check


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _req(t, dtype, ndim, name):
    if not isinstance(t, torch.Tensor):
        raise PttError("%s must be a tensor" % name)
    if not t.is_cuda:
        raise PttError("%s must be a CUDA tensor (there is no CPU path)" % name)
    if t.dtype != dtype:
        raise PttError("%s must be %s, got %s" % (name, dtype, t.dtype))
    if t.dim() != ndim:
        raise PttError("%s must have %d dimensions, got shape %s" % (name, ndim, tuple(t.shape)))
    if not t.is_contiguous():
        raise PttError("%s must be contiguous" % name)
    return t


def _req_strided(t, dtype, ndim, name):
    """_req without the contiguity requirement (the callee takes the strides)."""
    if not isinstance(t, torch.Tensor):
        raise PttError("%s must be a tensor" % name)
    if not t.is_cuda:
        raise PttError("%s must be a CUDA tensor (there is no CPU path)" % name)
    if t.dtype != dtype:
        raise PttError("%s must be %s, got %s" % (name, dtype, t.dtype))
    if t.dim() != ndim:
        raise PttError("%s must have %d dimensions, got shape %s" % (name, ndim, tuple(t.shape)))
    return t


def _same_device(*ts):
    dev = ts[0].device
    for t in ts[1:]:
        if t is not None and t.device != dev:
            raise PttError("tensors live on different devices")
    return dev


class _DeviceGuard:
    def __init__(self, dev):
        self.g = torch.cuda.device(dev)

    def __enter__(self):
        self.g.__enter__()

    def __exit__(self, *a):
        return self.g.__exit__(*a)


_F, _I = torch.float32, torch.int32

# The tcgen05 contractions split every fp32 operand into two fp16 halves (hi + lo, ~22 significant bits) with SATURATING
# conversions: |x| > 65504 would clamp silently.  Weights are checked once, at pack time (one device->host read; the
# pack calls synchronise anyway); anything beyond SPLIT_MAX -- far outside what a BatchNorm-folded PTT layer holds -- is
# refused instead of being computed wrongly.  Small magnitudes degrade gracefully: the lo half goes subnormal below
# |x| ~ 0.12, which bounds the ABSOLUTE representation error of an operand by 2^-25 (3e-8), not its relative error.
SPLIT_MAX = 3.0e4


def check_split_range(what, *tensors):
    worst = 0.0
    for t in tensors:
        if t is not None and t.numel():
            worst = max(worst, float(t.detach().abs().max()))
    if not worst < SPLIT_MAX:
        raise PttError("%s: largest magnitude %.3g is outside the fp16 hi/lo split range (|w| < %.0f) of the tensor-core "
                       "path; rescale the layer (fold the factor into the next layer or the BatchNorm scale)"
                       % (what, worst, SPLIT_MAX))



def _workspace(nbytes, dev):
    return torch.empty(max(int(nbytes), 16) // 4 + 4, dtype=_F, device=dev)


# ------------------------------------------------------------------------------------------------
# the pointnet2_ops `_ext` surface (pointnet2_utils.py:48-287)
# ------------------------------------------------------------------------------------------------
def furthest_point_sampling(xyz, npoint, return_new_xyz=False):
    _req(xyz, _F, 3, "xyz")
    B, N, three = xyz.shape
    npoint = int(npoint)
    if three != 3 or N < 1 or npoint < 0:
        raise PttError("furthest_point_sampling: xyz must be (B,N>=1,3) and npoint >= 0")
    with _DeviceGuard(xyz.device):
        idx = torch.empty(B, npoint, dtype=_I, device=xyz.device)
        new_xyz = torch.empty(B, npoint, 3, dtype=_F, device=xyz.device) if return_new_xyz else None
        L = _lib.lib()
        ws_bytes = L.ptt_furthest_point_sampling_workspace_bytes(B, N, npoint)
        ws = _workspace(ws_bytes, xyz.device) if ws_bytes else None
        check(L.ptt_furthest_point_sampling(_ptr(xyz), B, N, npoint, _ptr(idx), _ptr(new_xyz), _ptr(ws), ws_bytes,
                                            _stream()), "ptt_furthest_point_sampling")
    return (idx, new_xyz) if return_new_xyz else idx


def furthest_point_sampling_with_dist(dist, npoint):
    _req(dist, _F, 3, "dist")
    B, N, N2 = dist.shape
    npoint = int(npoint)
    if N != N2 or N < 1 or npoint < 0:
        raise PttError("furthest_point_sampling_with_dist: dist must be (B,N,N)")
    with _DeviceGuard(dist.device):
        idx = torch.empty(B, npoint, dtype=_I, device=dist.device)
        L = _lib.lib()
        ws_bytes = L.ptt_furthest_point_sampling_with_dist_workspace_bytes(B, N, npoint)
        ws = _workspace(ws_bytes, dist.device)
        check(L.ptt_furthest_point_sampling_with_dist(_ptr(dist), B, N, npoint, _ptr(idx), _ptr(ws), ws_bytes, _stream()),
              "ptt_furthest_point_sampling_with_dist")
    return idx


def gather_points(points, idx):
    _req(points, _F, 3, "points"), _req(idx, _I, 2, "idx")
    dev = _same_device(points, idx)
    B, C, N = points.shape
    M = idx.shape[1]
    if idx.shape[0] != B:
        raise PttError("gather_points: batch mismatch")
    with _DeviceGuard(dev):
        out = torch.empty(B, C, M, dtype=_F, device=dev)
        check(_lib.lib().ptt_gather_points(_ptr(points), _ptr(idx), B, C, N, M, _ptr(out), _stream()), "ptt_gather_points")
    return out


def gather_points_grad(grad_out, idx, n):
    _req(grad_out, _F, 3, "grad_out"), _req(idx, _I, 2, "idx")
    dev = _same_device(grad_out, idx)
    B, C, M = grad_out.shape
    if tuple(idx.shape) != (B, M):
        raise PttError("gather_points_grad: idx must be (B,M)")
    with _DeviceGuard(dev):
        out = torch.empty(B, C, int(n), dtype=_F, device=dev)
        check(_lib.lib().ptt_gather_points_grad(_ptr(grad_out), _ptr(idx), B, C, int(n), M, _ptr(out), _stream()),
              "ptt_gather_points_grad")
    return out


def ball_query(new_xyz, xyz, radius, nsample):
    _req(new_xyz, _F, 3, "new_xyz"), _req(xyz, _F, 3, "xyz")
    dev = _same_device(new_xyz, xyz)
    B, M, _ = new_xyz.shape
    N = xyz.shape[1]
    if xyz.shape[0] != B or xyz.shape[2] != 3 or new_xyz.shape[2] != 3 or N < 1:
        raise PttError("ball_query: new_xyz (B,M,3), xyz (B,N,3)")
    with _DeviceGuard(dev):
        idx = torch.empty(B, M, int(nsample), dtype=_I, device=dev)
        check(_lib.lib().ptt_ball_query(_ptr(new_xyz), _ptr(xyz), B, N, M, float(radius), int(nsample), _ptr(idx), _stream()),
              "ptt_ball_query")
    return idx


def ball_query_nested(xyz, samples, npoints, radii, nsamples):
    """The ball queries of every set-abstraction layer of a backbone branch in one launch (ptt_ball_query_nested):
    xyz (B,N,3), samples (B,npoints[0],3) = xyz in FPS order; level l: centres samples[:, :npoints[l]] against xyz
    (l = 0) or samples[:, :npoints[l-1]].  Returns [idx_l (B,npoints[l],nsamples[l]) int32]."""
    _req(xyz, _F, 3, "xyz"), _req(samples, _F, 3, "samples")
    dev = _same_device(xyz, samples)
    B, N, _ = xyz.shape
    L = len(npoints)
    if not (1 <= L <= 4) or len(radii) != L or len(nsamples) != L:
        raise PttError("ball_query_nested: 1..4 levels with one radius / nsample each")
    if samples.shape[0] != B or samples.shape[1] != npoints[0] or samples.shape[2] != 3 or xyz.shape[2] != 3:
        raise PttError("ball_query_nested: samples must be (B,npoints[0],3)")
    if any(npoints[l] > npoints[l - 1] for l in range(1, L)) or min(npoints) < 1 or min(nsamples) < 1:
        raise PttError("ball_query_nested: npoints must be non-increasing and positive")
    with _DeviceGuard(dev):
        outs = [torch.empty(B, int(npoints[l]), int(nsamples[l]), dtype=_I, device=dev) for l in range(L)]
        h_m = (ctypes.c_int * L)(*[int(v) for v in npoints])
        h_r = (ctypes.c_float * L)(*[float(v) for v in radii])
        h_ns = (ctypes.c_int * L)(*[int(v) for v in nsamples])
        h_out = (ctypes.c_void_p * L)(*[o.data_ptr() for o in outs])
        check(_lib.lib().ptt_ball_query_nested(_ptr(xyz), _ptr(samples), B, N, L, h_m, h_r, h_ns, h_out, _stream()),
              "ptt_ball_query_nested")
    return outs


def group_points(points, idx):
    _req(points, _F, 3, "points"), _req(idx, _I, 3, "idx")
    dev = _same_device(points, idx)
    B, C, N = points.shape
    _, M, K = idx.shape
    if idx.shape[0] != B:
        raise PttError("group_points: batch mismatch")
    with _DeviceGuard(dev):
        out = torch.empty(B, C, M, K, dtype=_F, device=dev)
        check(_lib.lib().ptt_group_points(_ptr(points), _ptr(idx), B, C, N, M, K, _ptr(out), _stream()), "ptt_group_points")
    return out


def group_points_grad(grad_out, idx, n):
    _req(grad_out, _F, 4, "grad_out"), _req(idx, _I, 3, "idx")
    dev = _same_device(grad_out, idx)
    B, C, M, K = grad_out.shape
    if tuple(idx.shape) != (B, M, K):
        raise PttError("group_points_grad: idx must be (B,M,K)")
    with _DeviceGuard(dev):
        out = torch.empty(B, C, int(n), dtype=_F, device=dev)
        check(_lib.lib().ptt_group_points_grad(_ptr(grad_out), _ptr(idx), B, C, int(n), M, K, _ptr(out), _stream()),
              "ptt_group_points_grad")
    return out


def three_nn(unknown, known):
    _req(unknown, _F, 3, "unknown"), _req(known, _F, 3, "known")
    dev = _same_device(unknown, known)
    B, n, _ = unknown.shape
    m = known.shape[1]
    with _DeviceGuard(dev):
        dist2 = torch.empty(B, n, 3, dtype=_F, device=dev)
        idx = torch.empty(B, n, 3, dtype=_I, device=dev)
        check(_lib.lib().ptt_three_nn(_ptr(unknown), _ptr(known), B, n, m, _ptr(dist2), _ptr(idx), _stream()), "ptt_three_nn")
    return dist2, idx


def three_interpolate(points, idx, weight):
    _req(points, _F, 3, "points"), _req(idx, _I, 3, "idx"), _req(weight, _F, 3, "weight")
    dev = _same_device(points, idx, weight)
    B, c, m = points.shape
    n = idx.shape[1]
    with _DeviceGuard(dev):
        out = torch.empty(B, c, n, dtype=_F, device=dev)
        check(_lib.lib().ptt_three_interpolate(_ptr(points), _ptr(idx), _ptr(weight), B, c, m, n, _ptr(out), _stream()),
              "ptt_three_interpolate")
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    _req(grad_out, _F, 3, "grad_out"), _req(idx, _I, 3, "idx"), _req(weight, _F, 3, "weight")
    dev = _same_device(grad_out, idx, weight)
    B, c, n = grad_out.shape
    with _DeviceGuard(dev):
        out = torch.empty(B, c, int(m), dtype=_F, device=dev)
        check(_lib.lib().ptt_three_interpolate_grad(_ptr(grad_out), _ptr(idx), _ptr(weight), B, c, n, int(m), _ptr(out),
                                                    _stream()), "ptt_three_interpolate_grad")
    return out


# ------------------------------------------------------------------------------------------------
# layout helpers, kNN, linear
# ------------------------------------------------------------------------------------------------
def cm_to_pm(src, ld=None):
    """(B,C,N) -> (B,N,ld) point-major, zero padded."""
    _req(src, _F, 3, "src")
    B, C, N = src.shape
    ld = C if ld is None else int(ld)
    with _DeviceGuard(src.device):
        dst = torch.empty(B, N, ld, dtype=_F, device=src.device)
        check(_lib.lib().ptt_cm_to_pm(_ptr(src), B, C, N, _ptr(dst), ld, _stream()), "ptt_cm_to_pm")
    return dst


def pm_to_cm(src, C=None):
    """(B,N,ld)[:, :, :C] -> (B,C,N) channel-major."""
    _req(src, _F, 3, "src")
    B, N, ld = src.shape
    C = ld if C is None else int(C)
    with _DeviceGuard(src.device):
        dst = torch.empty(B, C, N, dtype=_F, device=src.device)
        check(_lib.lib().ptt_pm_to_cm(_ptr(src), ld, B, C, N, _ptr(dst), _stream()), "ptt_pm_to_cm")
    return dst


def knn(xyz, k, out=None):
    _req(xyz, _F, 3, "xyz")
    B, n, _ = xyz.shape
    with _DeviceGuard(xyz.device):
        if out is None:
            out = torch.empty(B, n, int(k), dtype=_I, device=xyz.device)
        elif tuple(_req(out, _I, 3, "out").shape) != (B, n, int(k)):
            raise PttError("knn: out must be (B,n,k) int32")
        check(_lib.lib().ptt_knn(_ptr(xyz), B, n, int(k), _ptr(out), _stream()), "ptt_knn")
    return out


class PackedLinear:
    """nn.Linear parameters in the library's packed layout (ptt_linear_pack)."""

    def __init__(self, weight, bias=None, check_range=True):
        """weight (Cout, K): any strides (a transposed view packs without a copy)."""
        _req_strided(weight, _F, 2, "weight")
        self.cout, self.k = weight.shape
        self.has_bias = bias is not None
        if check_range:
            check_split_range("linear weight", weight)
        L = _lib.lib()
        with _DeviceGuard(weight.device):
            self.params = torch.empty(L.ptt_linear_params_floats(self.k, self.cout), dtype=_F, device=weight.device)
        self.repack(weight, bias)

    def repack(self, weight, bias=None):
        """Pack new values of the same shape into the existing image (no allocation, no synchronisation, no range check:
        the training path repacks every step)."""
        _req_strided(weight, _F, 2, "weight")
        if tuple(weight.shape) != (self.cout, self.k):
            raise PttError("repack: weight must be (%d,%d)" % (self.cout, self.k))
        self.has_bias = bias is not None
        with _DeviceGuard(weight.device):
            check(_lib.lib().ptt_linear_pack_strided(_ptr(weight), weight.stride(0), weight.stride(1), _ptr(bias), self.k, self.cout,
                                                     _ptr(self.params), _stream()), "ptt_linear_pack_strided")
        return self

    def __call__(self, x, relu=False, residual=None, in_affine=None, ld_out=None):
        """x (R, ldx >= K): the first K columns of every row are the input (rows padded to a multiple of 4 floats take
        the tensor-core path); residual (R, ldr >= Cout) is added after the activation.  in_affine = (ka, kb): the
        input is relu(ka[k] * x + kb[k]) applied on load (training path).  ld_out: row stride of the result (>= Cout;
        the padding columns are left unwritten)."""
        _req(x, _F, 2, "x")
        R, ldx = x.shape
        if ldx < self.k:
            raise PttError("linear: x has %d columns, weight expects %d" % (ldx, self.k))
        ldr = 0
        if residual is not None:
            _req(residual, _F, 2, "residual")
            ldr = residual.shape[1]
            if residual.shape[0] != R or ldr < self.cout:
                raise PttError("linear: residual must be (R, ld >= Cout)")
        ldy = self.cout if ld_out is None else int(ld_out)
        if ldy < self.cout:
            raise PttError("linear: ld_out < Cout")
        ka = kb = None
        if in_affine is not None:
            ka, kb = in_affine
            _req(ka, _F, 1, "ka"), _req(kb, _F, 1, "kb")
            if ka.numel() < self.k or kb.numel() < self.k:
                raise PttError("linear: in_affine vectors shorter than K")
        with _DeviceGuard(x.device):
            y = torch.empty(R, ldy, dtype=_F, device=x.device) if ldy == self.cout else torch.zeros(R, ldy, dtype=_F, device=x.device)
            check(_lib.lib().ptt_linear_fwd_ex(_ptr(x), ldx, R, self.k, _ptr(ka), _ptr(kb), _ptr(self.params), self.cout, int(relu),
                                               _ptr(residual), ldr, _ptr(y), ldy, _stream()), "ptt_linear_fwd_ex")
        return y


def linear_with_stats(lin, x, in_affine=None, want_stats=True, ld_out=None, zero_pad=True):
    """y = lin(f(x)) without activation / residual, plus (optionally) the column sums of y and y*y as a (2, Cout) float64
    tensor (ptt_linear_fwd_stats: bias-free layers over many rows take the weight-stationary persistent kernel).
    ld_out > Cout: rows are padded; zero_pad=False leaves the padding columns uninitialised (for consumers that never
    read them -- it saves a fill pass over the whole result)."""
    _req(x, _F, 2, "x")
    R, ldx = x.shape
    if ldx < lin.k:
        raise PttError("linear: x has %d columns, weight expects %d" % (ldx, lin.k))
    ka, kb = in_affine if in_affine is not None else (None, None)
    ldy = lin.cout if ld_out is None else int(ld_out)
    with _DeviceGuard(x.device):
        y = (torch.empty if (ldy == lin.cout or not zero_pad) else torch.zeros)(R, ldy, dtype=_F, device=x.device)
        sums = torch.empty(2, lin.cout, dtype=torch.float64, device=x.device) if want_stats else None
        check(_lib.lib().ptt_linear_fwd_stats(_ptr(x), ldx, R, lin.k, _ptr(ka), _ptr(kb), _ptr(lin.params), lin.cout,
                                              int(lin.has_bias), _ptr(y), ldy, _ptr(sums), _stream()), "ptt_linear_fwd_stats")
    return y, sums


class PackedConvStack:
    """pytorch_utils.Seq of 1x1 Conv1d layers (pytorch_utils.py:270-300; keys {i}.conv.weight (Cout,Cin,1), {i}.conv.bias,
    {i}.normlayer.bn.*) in eval mode: BatchNorm folded into the weights, ReLU after every layer but the last -- the
    heads' cla / vote / refine stacks and the similarity module's conv all end with activation=None
    (centroids_voting_head.py:14-25, box_voting_head.py:24-29, p2b_xcoor.py:20-24).  Rows are points."""

    def __init__(self, sd, eps=1e-5):
        self.layers = []
        i = 0
        while "%d.conv.weight" % i in sd:
            p = "%d." % i
            w = sd[p + "conv.weight"]
            w = w.reshape(w.shape[0], w.shape[1])
            b = sd.get(p + "conv.bias")
            if p + "normlayer.bn.weight" in sd:
                scale, shift = fold_batchnorm(sd[p + "normlayer.bn.weight"], sd[p + "normlayer.bn.bias"],
                                              sd[p + "normlayer.bn.running_mean"], sd[p + "normlayer.bn.running_var"], eps)
                w = w * scale[:, None]
                b = shift if b is None else shift + scale * b
            self.layers.append(PackedLinear(w.contiguous(), b.contiguous() if b is not None else None))
            i += 1
        if not self.layers:
            raise PttError("PackedConvStack: no layers in the state_dict")
        self.cout = self.layers[-1].cout

    def __call__(self, x, residual=None):
        n = len(self.layers)
        for i, lin in enumerate(self.layers):
            x = lin(x, relu=i < n - 1, residual=residual if i == n - 1 else None)
        return x


# ------------------------------------------------------------------------------------------------
# fused SA layer body and transformer block
# ------------------------------------------------------------------------------------------------
def fold_batchnorm(weight, bias, running_mean, running_var, eps=1e-5):
    """BatchNorm2d (eval) -> per-channel (scale, shift):  y = scale * x + shift."""
    scale = weight / torch.sqrt(running_var + eps)
    return scale.contiguous(), (bias - running_mean * scale).contiguous()


class PackedSAMlp:
    """SharedMLP of one SA layer (pytorch_utils.py:12-36) packed for ptt_sa_mlp_fwd.

    weights[l]: conv weight (Cout_l, Cin_l) with layer-0 input channels in the reference order
    [xyz(3) | feats(C)]; scales/shifts: folded BatchNorm (or None)."""

    def __init__(self, weights, scales, shifts):
        self.n_layers = len(weights)
        ws = [_req(w.reshape(w.shape[0], w.shape[1]).contiguous(), _F, 2, "weight") for w in weights]
        dims = [ws[0].shape[1]] + [w.shape[0] for w in ws]
        self.C = dims[0] - 3
        self.dims = dims
        self.h_dims = (ctypes.c_int * len(dims))(*dims)
        dev = ws[0].device
        L = _lib.lib()
        n = L.ptt_sa_params_floats(self.C, self.n_layers, self.h_dims)
        if n == 0:
            raise PttError("ptt_sa_params_floats rejected dims %s" % (dims,))

        def arr(ts):
            return (ctypes.c_void_p * self.n_layers)(*[t.data_ptr() if t is not None else None for t in ts])

        scales = [s.contiguous() if s is not None else None for s in (scales or [None] * self.n_layers)]
        shifts = [s.contiguous() if s is not None else None for s in (shifts or [None] * self.n_layers)]
        check_split_range("SA layer weight (BatchNorm scale folded)",
                          *[w * sc[:, None] if sc is not None else w for w, sc in zip(ws, scales)])
        with _DeviceGuard(dev):
            self.params = torch.empty(n, dtype=_F, device=dev)
            check(L.ptt_sa_pack_params(self.C, self.n_layers, self.h_dims, arr(ws), arr(scales), arr(shifts),
                                       _ptr(self.params), _stream()), "ptt_sa_pack_params")
            torch.cuda.current_stream().synchronize()   # ws/scales/shifts temporaries may die after return
        self.cout = dims[-1]

    def workspace_bytes(self, B, N, M, ns):
        return _lib.lib().ptt_sa_mlp_workspace_bytes(B, N, M, ns, self.C, self.n_layers, self.h_dims)


def sa_mlp_fwd(packed, xyz, feats_pm, new_xyz, idx, radius, normalize_xyz, want_pm=True, want_cm=True, workspace=None):
    """xyz (B,N,3), feats_pm (B,N,ldf)|None, new_xyz (B,M,3), idx (B,M,ns) -> out_pm (B,M,Cout), out_cm (B,Cout,M)."""
    _req(xyz, _F, 3, "xyz"), _req(new_xyz, _F, 3, "new_xyz"), _req(idx, _I, 3, "idx")
    dev = _same_device(xyz, new_xyz, idx, feats_pm)
    B, N, _ = xyz.shape
    _, M, ns = idx.shape
    ldf = 0
    if packed.C > 0:
        _req(feats_pm, _F, 3, "feats_pm")
        ldf = feats_pm.shape[2]
        if feats_pm.shape[0] != B or feats_pm.shape[1] != N or ldf < packed.C:
            raise PttError("sa_mlp_fwd: feats_pm must be (B,N,ld>=C)")
    L = _lib.lib()
    with _DeviceGuard(dev):
        out_pm = torch.empty(B, M, packed.cout, dtype=_F, device=dev) if want_pm else None
        out_cm = torch.empty(B, packed.cout, M, dtype=_F, device=dev) if want_cm else None
        ws_bytes = packed.workspace_bytes(B, N, M, ns)
        ws = workspace if workspace is not None else _workspace(ws_bytes, dev)
        check(L.ptt_sa_mlp_fwd(_ptr(xyz), _ptr(feats_pm) if packed.C > 0 else None, ldf, _ptr(new_xyz), _ptr(idx), B, N, M,
                               ns, packed.C, float(radius), int(bool(normalize_xyz)), packed.n_layers, packed.h_dims,
                               _ptr(packed.params), _ptr(out_pm), packed.cout, _ptr(out_cm), _ptr(ws),
                               ws.numel() * 4, _stream()), "ptt_sa_mlp_fwd")
    return out_pm, out_cm


class PackedCosineFusion:
    """CosineSimAug (similarity_modules/p2b_xcoor.py:9-46), eval mode, from its state_dict (keys mlp.layer{i}.*, conv.{i}.*).

    The (1 + 3 + f)-channel fusion rows only depend on the search point through ONE channel (the cosine similarity), so
    the SharedMLP + max over the templates is a set-abstraction layer over the templates (ptt_cosine_fusion_fwd)."""

    def __init__(self, sd, eps=1e-5):
        ws, scales, shifts = [], [], []
        i = 0
        while "mlp.layer%d.conv.weight" % i in sd:
            p = "mlp.layer%d." % i
            w = sd[p + "conv.weight"]
            w = w.reshape(w.shape[0], w.shape[1])
            if i == 0:      # reference columns [sim | xyz_t(3) | feats_t(f)] -> SA order [rel(3) = (sim, 0, 0) | point features]
                w = torch.cat([w[:, :1], torch.zeros_like(w[:, :2]), w[:, 1:]], dim=1)
            ws.append(w.contiguous())
            if p + "normlayer.bn.weight" in sd:
                sc, sh = fold_batchnorm(sd[p + "normlayer.bn.weight"], sd[p + "normlayer.bn.bias"],
                                        sd[p + "normlayer.bn.running_mean"], sd[p + "normlayer.bn.running_var"], eps)
                if p + "conv.bias" in sd:
                    sh = sh + sd[p + "conv.bias"] * sc
            else:
                sc, sh = None, sd.get(p + "conv.bias")
            scales.append(sc)
            shifts.append(sh)
            i += 1
        self.mlp = PackedSAMlp(ws, scales, shifts)
        self.f = self.mlp.C - 3
        n = len("conv.")
        self.conv = PackedConvStack({k[n:]: v for k, v in sd.items() if k.startswith("conv.")}, eps)

    def workspace_bytes(self, B, n1, n2):
        return _lib.lib().ptt_cosine_fusion_workspace_bytes(B, n1, n2, self.f, self.mlp.n_layers, self.mlp.h_dims)

    def __call__(self, search_feats_pm, template_feats_pm, template_xyz, workspace=None):
        """search_feats_pm (B,n2,f), template_feats_pm (B,n1,f), template_xyz (B,n1,3) -> cosine_feats_pm (B,n2,Cout)."""
        _req(search_feats_pm, _F, 3, "search_feats_pm"), _req(template_feats_pm, _F, 3, "template_feats_pm")
        _req(template_xyz, _F, 3, "template_xyz")
        dev = _same_device(search_feats_pm, template_feats_pm, template_xyz)
        B, n2, lds = search_feats_pm.shape
        _, n1, ldt = template_feats_pm.shape
        if lds < self.f or ldt < self.f or template_feats_pm.shape[0] != B or tuple(template_xyz.shape) != (B, n1, 3):
            raise PttError("cosine fusion: inconsistent shapes")
        L = _lib.lib()
        with _DeviceGuard(dev):
            fused = torch.empty(B, n2, self.mlp.cout, dtype=_F, device=dev)
            ws_bytes = self.workspace_bytes(B, n1, n2)
            ws = workspace if workspace is not None else _workspace(ws_bytes, dev)
            check(L.ptt_cosine_fusion_fwd(_ptr(search_feats_pm), lds, _ptr(template_feats_pm), ldt, _ptr(template_xyz), B, n1,
                                          n2, self.f, self.mlp.n_layers, self.mlp.h_dims, _ptr(self.mlp.params), _ptr(fused),
                                          self.mlp.cout, _ptr(ws), ws.numel() * 4, _stream()), "ptt_cosine_fusion_fwd")
        return self.conv(fused.reshape(B * n2, self.mlp.cout)).reshape(B, n2, -1)


TRANSFORMER_KEYS = ("fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias", "fc_delta.0.weight", "fc_delta.0.bias",
                    "fc_delta.2.weight", "fc_delta.2.bias", "fc_gamma.0.weight", "fc_gamma.0.bias",
                    "fc_gamma.2.weight", "fc_gamma.2.bias", "w_qs.weight", "w_ks.weight", "w_vs.weight")


class PackedTransformer:
    """TransformerBlock parameters (state_dict keys of variants.py:129-147) packed for ptt_transformer_block_fwd."""

    def __init__(self, sd, k, variant=0, check_range=True):
        ts = [_req(sd[key].contiguous(), _F, sd[key].dim(), key) for key in TRANSFORMER_KEYS]
        if check_range:
            check_split_range("transformer block weight", *[t for t in ts if t.dim() == 2])
        self.d_model, self.d_points = ts[0].shape
        self.k = int(k)
        self.variant = int(variant)
        dev = ts[0].device
        L = _lib.lib()
        with _DeviceGuard(dev):
            self.params = torch.empty(L.ptt_transformer_params_floats(self.d_points, self.d_model), dtype=_F, device=dev)
            check(L.ptt_transformer_pack_params(self.d_points, self.d_model, *[_ptr(t) for t in ts], _ptr(self.params),
                                                _stream()), "ptt_transformer_pack_params")
            if check_range:
                torch.cuda.current_stream().synchronize()     # the .contiguous() temporaries may die after return
            else:
                self._keep = ts                               # training path: no sync; the sources stay referenced

    def workspace_bytes(self, B, n):
        return _lib.lib().ptt_transformer_block_workspace_bytes(B, n, self.k, self.d_points, self.d_model)


def transformer_block_fwd(packed, xyz, features, knn_idx=None, want_attn=False, workspace=None):
    """xyz (B,n,3), features (B,n,d_points) -> out (B,n,d_points) [, attn (B,n,k,d_model)]."""
    _req(xyz, _F, 3, "xyz"), _req(features, _F, 3, "features")
    dev = _same_device(xyz, features, knn_idx)
    B, n, _ = xyz.shape
    if tuple(features.shape) != (B, n, packed.d_points):
        raise PttError("transformer_block_fwd: features must be (B,n,%d)" % packed.d_points)
    if knn_idx is not None:
        _req(knn_idx, _I, 3, "knn_idx")
    L = _lib.lib()
    with _DeviceGuard(dev):
        out = torch.empty(B, n, packed.d_points, dtype=_F, device=dev)
        attn = torch.empty(B, n, packed.k, packed.d_model, dtype=_F, device=dev) if want_attn else None
        ws_bytes = packed.workspace_bytes(B, n)
        ws = workspace if workspace is not None else _workspace(ws_bytes, dev)
        check(L.ptt_transformer_block_fwd(_ptr(xyz), _ptr(features), B, n, packed.k, packed.d_points, packed.d_model,
                                          packed.variant, _ptr(packed.params), _ptr(knn_idx), _ptr(out), _ptr(attn),
                                          _ptr(ws), ws.numel() * 4, _stream()), "ptt_transformer_block_fwd")
    return (out, attn) if want_attn else out


def transformer_block_fwd_ex(packed, xyz, features, q_features=None, flags=0, divisor=0.0, knn_idx=None, pair_scalar=None,
                             pair_vec=None, want_attn=False, workspace=None):
    """ptt_transformer_block_fwd_ex: flags bit 0 = Offset, bit 1 = raw (returns res (B,n,d_model))."""
    _req(xyz, _F, 3, "xyz"), _req(features, _F, 3, "features")
    dev = _same_device(xyz, features, knn_idx, q_features, pair_scalar, pair_vec)
    B, n, _ = xyz.shape
    if tuple(features.shape) != (B, n, packed.d_points):
        raise PttError("transformer_block_fwd_ex: features must be (B,n,%d)" % packed.d_points)
    if q_features is not None and tuple(_req(q_features, _F, 3, "q_features").shape) != tuple(features.shape):
        raise PttError("transformer_block_fwd_ex: q_features must have the shape of features")
    if knn_idx is not None:
        _req(knn_idx, _I, 3, "knn_idx")
    if (pair_scalar is None) != (pair_vec is None):
        raise PttError("transformer_block_fwd_ex: pair_scalar and pair_vec come together")
    if pair_scalar is not None:
        _req(pair_scalar, _F, 3, "pair_scalar"), _req(pair_vec, _F, 1, "pair_vec")
    L = _lib.lib()
    with _DeviceGuard(dev):
        out = torch.empty(B, n, packed.d_model if flags & 2 else packed.d_points, dtype=_F, device=dev)
        attn = torch.empty(B, n, packed.k, packed.d_model, dtype=_F, device=dev) if want_attn else None
        ws = workspace if workspace is not None else _workspace(packed.workspace_bytes(B, n), dev)
        check(L.ptt_transformer_block_fwd_ex(_ptr(xyz), _ptr(features), _ptr(q_features), B, n, packed.k, packed.d_points,
                                             packed.d_model, int(flags), float(divisor), _ptr(packed.params), _ptr(knn_idx),
                                             _ptr(pair_scalar), _ptr(pair_vec), _ptr(out), _ptr(attn), _ptr(ws),
                                             ws.numel() * 4, _stream()), "ptt_transformer_block_fwd_ex")
    return (out, attn) if want_attn else out


def pair_cosine(q, kmat, knn_idx):
    """q, kmat (B,n,d), knn_idx (B,n,k) int32 -> (B,n,k) cosine similarities."""
    _req(q, _F, 3, "q"), _req(kmat, _F, 3, "kmat"), _req(knn_idx, _I, 3, "knn_idx")
    dev = _same_device(q, kmat, knn_idx)
    B, n, d = q.shape
    k = knn_idx.shape[2]
    with _DeviceGuard(dev):
        sim = torch.empty(B, n, k, dtype=_F, device=dev)
        check(_lib.lib().ptt_pair_cosine(_ptr(q), d, _ptr(kmat), kmat.shape[2], _ptr(knn_idx), B, n, k, d, _ptr(sim), _stream()),
              "ptt_pair_cosine")
    return sim


def layer_norm(x, gamma, beta, eps=1e-5, residual=None):
    """x (R,C) -> LayerNorm over C (* gamma + beta) (+ residual (R,C))."""
    _req(x, _F, 2, "x")
    R, C = x.shape
    with _DeviceGuard(x.device):
        y = torch.empty_like(x)
        check(_lib.lib().ptt_layer_norm_fwd(_ptr(x), C, R, C, _ptr(gamma), _ptr(beta), float(eps), _ptr(residual),
                                            C if residual is not None else 0, _ptr(y), C, _stream()), "ptt_layer_norm_fwd")
    return y


def token_softmax_gate(logits, other, divisor, want_attn=False):
    """logits, other (B,n,C): softmax over the n tokens per (cloud, channel), times other."""
    _req(logits, _F, 3, "logits"), _req(other, _F, 3, "other")
    B, n, C = logits.shape
    with _DeviceGuard(logits.device):
        out = torch.empty_like(logits)
        attn = torch.empty_like(logits) if want_attn else None
        check(_lib.lib().ptt_token_softmax_gate(_ptr(logits), C, _ptr(other), C, B, n, C, float(divisor), _ptr(out), C,
                                                _ptr(attn), _stream()), "ptt_token_softmax_gate")
    return (out, attn) if want_attn else out


STD_KEYS = ("fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias", "fc_delta.0.weight", "fc_delta.0.bias",
            "fc_delta.2.weight", "fc_delta.2.bias", "w_qs.weight", "w_ks.weight", "w_vs.weight")


class PackedTransformerSTD:
    """TransformerBlockSTD parameters (variants.py:13-27) packed for ptt_transformer_std_fwd."""

    def __init__(self, sd):
        ts = [_req(sd[key].contiguous(), _F, sd[key].dim(), key) for key in STD_KEYS]
        check_split_range("transformer block weight", *[t for t in ts if t.dim() == 2])
        self.d_model, self.d_points = ts[0].shape
        dev = ts[0].device
        L = _lib.lib()
        with _DeviceGuard(dev):
            self.params = torch.empty(L.ptt_transformer_std_params_floats(self.d_points, self.d_model), dtype=_F, device=dev)
            check(L.ptt_transformer_std_pack_params(self.d_points, self.d_model, *[_ptr(t) for t in ts], _ptr(self.params),
                                                    _stream()), "ptt_transformer_std_pack_params")
            torch.cuda.current_stream().synchronize()


def transformer_std_fwd(packed, xyz, features, want_attn=True, workspace=None):
    """xyz (B,n,3), features (B,n,d_points) -> out (B,n,d_points) [, attn (B,n,n)]."""
    _req(xyz, _F, 3, "xyz"), _req(features, _F, 3, "features")
    dev = _same_device(xyz, features)
    B, n, _ = xyz.shape
    if tuple(features.shape) != (B, n, packed.d_points):
        raise PttError("transformer_std_fwd: features must be (B,n,%d)" % packed.d_points)
    L = _lib.lib()
    with _DeviceGuard(dev):
        out = torch.empty(B, n, packed.d_points, dtype=_F, device=dev)
        attn = torch.empty(B, n, n, dtype=_F, device=dev) if want_attn else None
        ws_bytes = L.ptt_transformer_std_workspace_bytes(B, n, packed.d_points, packed.d_model)
        ws = workspace if workspace is not None else _workspace(ws_bytes, dev)
        check(L.ptt_transformer_std_fwd(_ptr(xyz), _ptr(features), B, n, packed.d_points, packed.d_model, _ptr(packed.params),
                                        _ptr(out), _ptr(attn), _ptr(ws), ws.numel() * 4, _stream()), "ptt_transformer_std_fwd")
    return (out, attn) if want_attn else out


# ------------------------------------------------------------------------------------------------
# N3: per-frame pre / post-processing of the tracking loop (csrc/tracking.cu)
# ------------------------------------------------------------------------------------------------
_F64 = torch.float64


def mt19937_stream(n, seed=1, device=None):
    """First n 32-bit outputs of MT19937 seeded like np.random.seed(seed), as an int32-typed (bit pattern) tensor."""
    host = torch.empty(int(n), dtype=_I)
    check(_lib.lib().ptt_mt19937_stream(int(seed), int(n), ctypes.c_void_p(host.data_ptr())), "ptt_mt19937_stream")
    return host.to(device) if device is not None else host


def track_crop(sources, offset, scale, search, out, out_counts):
    """sources: 1 or 2 tuples (points (T,cap,3) f32, counts (T) i32, boxes (T,15) f64 | None = precropped) -> out
    (T,cap_out,3), out_counts (T) filled in place (ptt_track_crop)."""
    n = len(sources)
    if not 1 <= n <= 2:
        raise PttError("track_crop: 1 or 2 sources")
    T = out.shape[0]
    _req(out, _F, 3, "out"), _req(out_counts, _I, 1, "out_counts")
    need = 0
    for pts, cnt, box in sources:
        _req(pts, _F, 3, "points"), _req(cnt, _I, 1, "counts")
        if box is not None:
            _req(box, _F64, 2, "boxes")
            if tuple(box.shape) != (T, 15):
                raise PttError("track_crop: boxes must be (T,15) float64")
        if pts.shape[0] != T or pts.shape[2] != 3 or cnt.shape[0] != T:
            raise PttError("track_crop: points (T,cap,3), counts (T)")
        need += pts.shape[1]
    if out.shape[1] < need or out.shape[2] != 3 or out_counts.shape[0] != T:
        raise PttError("track_crop: out must hold the capacities of all sources (%d rows)" % need)
    arr = lambda xs: (ctypes.c_void_p * n)(*[x.data_ptr() if x is not None else None for x in xs])
    with _DeviceGuard(out.device):
        check(_lib.lib().ptt_track_crop(T, n, arr([s[0] for s in sources]), arr([s[1] for s in sources]),
                                        arr([s[2] for s in sources]), (ctypes.c_int * n)(*[s[0].shape[1] for s in sources]),
                                        (ctypes.c_int * n)(*[int(s[2] is None) for s in sources]), float(offset), float(scale),
                                        int(bool(search)), _ptr(out), out.shape[1], _ptr(out_counts), _stream()), "ptt_track_crop")


def track_regularize(points, counts, size, mt, mt_pos, out):
    """points (T,cap,3), counts (T) -> out (T,size,3) (ptt_track_regularize); mt_pos (T) i32 updated in place."""
    _req(points, _F, 3, "points"), _req(counts, _I, 1, "counts"), _req(mt, _I, 1, "mt"), _req(mt_pos, _I, 1, "mt_pos")
    _req(out, _F, 3, "out")
    T, cap, _ = points.shape
    if tuple(out.shape) != (T, int(size), 3):
        raise PttError("track_regularize: out must be (T,size,3)")
    with _DeviceGuard(points.device):
        check(_lib.lib().ptt_track_regularize(_ptr(points), _ptr(counts), T, cap, int(size), _ptr(mt), mt.numel(), _ptr(mt_pos),
                                              _ptr(out), _stream()), "ptt_track_regularize")


def track_update(best_box, state, use_z, mt, mt_pos, results, frame_idx):
    """best_box (T,>=4) f32, state (T,15) f64 in place, results (F,T,15) f64, frame_idx (1,) i32 (ptt_track_update)."""
    _req(best_box, _F, 2, "best_box"), _req(state, _F64, 2, "state"), _req(results, _F64, 3, "results")
    _req(frame_idx, _I, 1, "frame_idx"), _req(mt, _I, 1, "mt"), _req(mt_pos, _I, 1, "mt_pos")
    T = state.shape[0]
    if best_box.shape[0] != T or best_box.shape[1] < 4 or state.shape[1] != 15 or tuple(results.shape[1:]) != (T, 15):
        raise PttError("track_update: inconsistent shapes")
    with _DeviceGuard(state.device):
        check(_lib.lib().ptt_track_update(_ptr(best_box), best_box.shape[1], _ptr(state), T, int(bool(use_z)), _ptr(mt),
                                          mt.numel(), _ptr(mt_pos), _ptr(results), results.shape[0], _ptr(frame_idx), _stream()),
              "ptt_track_update")
