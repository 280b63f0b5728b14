"""GPU: the native training path (ptt_b200/train_ops.py; csrc/train_ops.cu, tc_wgrad.cu, tr_train.cu, tc_gemm.cu's operand
transform) against torch autograd over the reference's decomposition (fp32, TF32 off): forward values, every gradient,
the BatchNorm running statistics.

Tolerances.  Kernel-level checks (one op against torch on identical inputs): 1e-4 .. 2e-5 of the tensor's scale.
Module-level gradients: a ReLU network's gradient is discontinuous where a pre-activation crosses zero, and two fp32
implementations of the same forward differ by ~1e-6 there, so among the millions of activations of a full-size layer a
handful get the other sub-gradient ("mask flips": measured 7 of 4.2 M at SA2).  Each flip moves one element of dy by
O(1), i.e. a weight-gradient row by ~1e-2 of the gradient's scale -- both results are exact gradients of their own
forward pass.  Module-level gradients are therefore compared in flip-robust metrics (cosine similarity and the
fraction of entries outside 1e-3 of the scale) at full size, and in the max norm on layers small enough for flips not
to occur."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import t
from ptt_b200 import modules, ops, synth, train_ops
from test_oracle_golden import sa_state_dict

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


@pytest.fixture(autouse=True)
def _true_fp32():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def close(a, b, tol, what=""):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    scale = max(float(b.abs().max()), 1e-6)
    err = float((a - b).abs().max())
    assert err <= tol * scale, "%s: max |diff| %.3e vs scale %.3e (tol %.1e)" % (what, err, scale, tol)


@pytest.mark.parametrize("R,M,N", [(5000, 128, 131), (64, 64, 3), (12345, 256, 259), (4096, 256, 128), (70000, 512, 512),
                                   (300, 100, 36), (8200, 128, 64), (9000, 256, 64), (3000, 512, 3), (2000, 384, 192)])
def test_linear_wgrad_vs_torch(R, M, N):
    rs = np.random.RandomState(R + M)
    ldy, ldx = (M + 3) // 4 * 4, (N + 3) // 4 * 4
    dy = torch.from_numpy(rs.standard_normal((R, ldy)).astype(np.float32)).to(DEV)
    x = torch.from_numpy(rs.standard_normal((R, ldx)).astype(np.float32)).to(DEV)
    want = dy[:, :M].double().t() @ x[:, :N].double()
    got = train_ops.linear_wgrad(dy, x, M, N)
    close(got, want, 2e-5, "plain")
    ka = torch.from_numpy(rs.uniform(0.5, 1.5, N).astype(np.float32)).to(DEV)
    kb = torch.from_numpy(rs.normal(0, 0.5, N).astype(np.float32)).to(DEV)
    want = dy[:, :M].double().t() @ torch.relu(x[:, :N] * ka + kb).double()
    got, db = train_ops.linear_wgrad(dy, x, M, N, (ka, kb), want_bias=True)
    close(got, want, 2e-5, "affine + relu on load")
    assert db.shape == (M,)
    close(db, dy[:, :M].double().sum(0), 2e-5, "bias gradient (column sums of dy) from the same kernel")


def test_packed_linear_from_a_transposed_view_needs_no_copy():
    """ptt_linear_pack_strided: the input-gradient contraction packs W^T straight from the (Cout, K) parameter."""
    rs = np.random.RandomState(5)
    for cout, k in ((128, 131), (64, 6), (512, 512), (100, 36)):
        w = torch.from_numpy(rs.standard_normal((cout, k)).astype(np.float32)).to(DEV)
        a = ops.PackedLinear(w.t(), None)                 # (K, Cout) view with strides (1, K)
        b = ops.PackedLinear(w.t().contiguous(), None)
        assert not w.t().is_contiguous() and torch.equal(a.params, b.params)
    bias = torch.from_numpy(rs.standard_normal(64).astype(np.float32)).to(DEV)
    w = torch.from_numpy(rs.standard_normal((64, 20)).astype(np.float32)).to(DEV)
    lin = ops.PackedLinear(w, bias)
    ldw = 64
    assert torch.equal(lin.params[:20 * ldw].view(20, ldw), w.t()) and torch.equal(lin.params[20 * ldw: 21 * ldw], bias)


def test_linear_fwd_operand_transform_and_padding():
    rs = np.random.RandomState(3)
    for R, K, Cout in ((1000, 128, 256), (333, 64, 131), (129, 260, 256)):
        ld = (K + 3) // 4 * 4
        x = torch.from_numpy(rs.standard_normal((R, ld)).astype(np.float32)).to(DEV)
        w = torch.from_numpy((rs.standard_normal((Cout, K)) / np.sqrt(K)).astype(np.float32)).to(DEV)
        ka = torch.from_numpy(rs.uniform(0.5, 1.5, K).astype(np.float32)).to(DEV)
        kb = torch.from_numpy(rs.normal(0, 0.5, K).astype(np.float32)).to(DEV)
        lin = ops.PackedLinear(w)
        want = torch.relu(x[:, :K] * ka + kb).double() @ w.double().t()
        got = lin(x, in_affine=(ka, kb), ld_out=(Cout + 3) // 4 * 4)
        close(got[:, :Cout], want, 2e-5, "fwd_ex %s" % ((R, K, Cout),))
        w2 = w * 0.5
        close(lin.repack(w2)(x), x[:, :K].double() @ w2.double().t(), 2e-5, "repack")


def test_two_phase_batchnorm_and_pool_vs_torch():
    rs = np.random.RandomState(5)
    for groups, ns, C in ((500, 32, 128), (77, 16, 256), (100, 8, 48)):
        R = groups * ns
        y = torch.from_numpy(rs.standard_normal((R, C)).astype(np.float32) * 2 + 0.3).to(DEV)
        gamma = torch.from_numpy(rs.uniform(0.5, 1.5, C).astype(np.float32)).to(DEV)
        beta = torch.from_numpy(rs.normal(0, 0.3, C).astype(np.float32)).to(DEV)
        rm, rv = torch.zeros(C, device=DEV), torch.ones(C, device=DEV)
        rm2, rv2 = rm.clone(), rv.clone()
        ka, kb, mean, rstd = train_ops.bn_train_finalize(train_ops.col_stats(y, C), R, gamma, beta, 1e-5, 0.1, rm, rv)
        yt = y.clone().requires_grad_(True)
        z = torch.relu(F.batch_norm(yt.t().reshape(1, C, R), rm2, rv2, gamma, beta, training=True, momentum=0.1, eps=1e-5))
        close(torch.relu(y * ka + kb), z[0].t(), 1e-5, "normalised")
        close(rm, rm2, 1e-5, "running_mean"), close(rv, rv2, 1e-5, "running_var")
        pooled, arg = train_ops.bn_relu_maxpool(y, groups, ns, C, ka, kb)
        zt = z[0].t().reshape(groups, ns, C)
        close(pooled, zt.max(1)[0], 1e-5, "max pool")
        # backward through pool + relu + batch norm
        dout = torch.from_numpy(rs.standard_normal((groups, C)).astype(np.float32)).to(DEV)
        zt.max(1)[0].backward(dout)
        dy, sums, _ = train_ops.bn_relu_bwd(dout, arg, ns, y, C, ka, kb, mean, rstd, gamma)
        close(dy, yt.grad, 1e-4, "dy (pooled)")
        # dense dz
        yt.grad = None
        dz = torch.from_numpy(rs.standard_normal((R, C)).astype(np.float32)).to(DEV)
        gam = gamma.clone().requires_grad_(True)
        bet = beta.clone().requires_grad_(True)
        z2 = torch.relu(F.batch_norm(yt.t().reshape(1, C, R), None, None, gam, bet, training=True, eps=1e-5))
        z2[0].t().backward(dz)
        dy, sums, dparam = train_ops.bn_relu_bwd(dz, None, 1, y, C, ka, kb, mean, rstd, gamma)
        close(dy, yt.grad, 1e-4, "dy (dense)")
        close(sums[1], gam.grad, 1e-4, "d gamma"), close(sums[0], bet.grad, 1e-4, "d beta")
        assert dparam.dtype == torch.float32 and torch.equal(dparam, sums.float()), "fp32 parameter gradients = the rounded sums"


def close_grad(a, b, what="", cos_tol=3e-4, l2_tol=2.5e-2):
    """Flip-robust comparison (see the module docstring): direction and relative L2 distance of the two gradients."""
    a, b = a.detach().double().cpu().reshape(-1), b.detach().double().cpu().reshape(-1)
    cos = float((a * b).sum() / (a.norm() * b.norm()).clamp_min(1e-30))
    l2 = float((a - b).norm() / b.norm().clamp_min(1e-30))
    assert cos >= 1 - cos_tol and l2 <= l2_tol, "%s: cosine %.6f, relative L2 distance %.2e" % (what, cos, l2)


SA_TRAIN_CASES = [  # (B, N, C_in, mlp, npoint, radius, ns, xyz requires grad)
    (4, 512, 128, [128, 128, 128, 256], 256, 0.5, 32, False),        # ptt.yaml SA2
    (3, 1024, 0, [0, 64, 64, 128], 512, 0.3, 32, False),             # SA1: xyz only
    (4, 128, 257, [257, 256, 256, 256], 64, 0.3, 16, True),          # the box head's vote aggregation (xyz = votes: grad)
    (2, 100, 16, [16, 32, 48], 40, 0.6, 8, True),
]


@pytest.mark.parametrize("case", SA_TRAIN_CASES)
def test_sa_module_native_training_vs_torch_decomposition(case):
    B, N, cin, mlp, npoint, radius, ns, xyz_grad = case
    sd = {k: t(v) for k, v in synth.fill_state_dict(sa_state_dict(mlp), seed=500 + N).items()}
    xyz = torch.from_numpy(synth.make_clouds(B, N, 600 + N, "dense", role="template" if N <= 512 else "search")).to(DEV)
    feats = torch.from_numpy(synth.features((B, cin, N), seed=601 + N)).to(DEV) if cin else None
    outs = []
    for native in (True, False):
        mod = modules.PointnetSAModuleVotes(mlp=list(mlp), radius=radius, nsample=ns, normalize_xyz=True, sample_method="fps")
        mod.load_state_dict(sd)
        mod = mod.to(DEV).train()
        mod.native_train = native
        x = xyz.clone().requires_grad_(xyz_grad)
        f = feats.clone().requires_grad_(True) if cin else None
        new_xyz, new_feats, inds = mod(x, f, npoint)
        w = torch.from_numpy(synth.features(tuple(new_feats.shape), seed=7)).to(DEV)
        ((new_feats * w).sum() + (new_xyz.sum() if xyz_grad else 0.0)).backward()
        outs.append(dict(feats=new_feats, inds=inds, fg=f.grad if cin else None, xg=x.grad if xyz_grad else None,
                         params={k: p.grad for k, p in mod.named_parameters()},
                         buffers={k: b.clone() for k, b in mod.named_buffers()}))
    a, b = outs
    assert torch.equal(a["inds"], b["inds"])
    close(a["feats"], b["feats"], 1e-4, "forward")
    small = B * npoint * ns * max(mlp) < 200000          # few enough activations for ReLU mask flips not to occur
    cmp = (lambda x, y, w: close(x, y, 1e-3, w)) if small else close_grad
    if cin:
        cmp(a["fg"], b["fg"], "d features")
    if xyz_grad:
        cmp(a["xg"], b["xg"], "d xyz")
    for k in b["params"]:
        cmp(a["params"][k], b["params"][k], "d " + k)
    for k in b["buffers"]:
        close(a["buffers"][k], b["buffers"][k], 1e-5, k)


@pytest.mark.parametrize("case", [(4, 128, 256, 512, 16), (3, 64, 256, 512, 16), (2, 48, 32, 64, 8), (2, 40, 64, 128, 4)])
def test_transformer_block_native_training_vs_torch_decomposition(case):
    from test_oracle_golden import transformer_state_dict
    B, n, dp, dm, k = case
    sd = {kk: t(v) for kk, v in synth.fill_state_dict(transformer_state_dict("TransformerBlock", dp, dm), seed=700 + n).items()}
    xyz = torch.from_numpy(synth.make_clouds(B, n, 701 + n, "dense", role="template")).to(DEV)
    feats = torch.from_numpy(synth.features((B, n, dp), seed=702 + n)).to(DEV)
    w = torch.from_numpy(synth.features((B, n, dp), seed=703 + n)).to(DEV)
    outs = []
    for native in (True, False):
        mod = modules.TransformerBlock(dp, dm, k)
        mod.load_state_dict(sd)
        mod = mod.to(DEV).train()
        mod.native_train = native
        f = feats.clone().requires_grad_(True)
        res, attn = mod(xyz, f)
        (res * w).sum().backward()
        outs.append(dict(res=res, attn=attn, fg=f.grad, params={kk: p.grad for kk, p in mod.named_parameters()}))
    a, b = outs
    close(a["res"], b["res"], 1e-4, "forward")
    close(a["attn"], b["attn"], 1e-4, "attn")
    small = B * n * k * dm < 200000
    cmp = (lambda x, y, what: close(x, y, 1e-3, what)) if small else close_grad
    cmp(a["fg"], b["fg"], "d features")
    for kk in b["params"]:
        if kk == "fc_gamma.2.bias":          # constant over a token's neighbours: cancels in the softmax, gradient == 0
            assert float(a["params"][kk].abs().max()) < 1e-3 and float(b["params"][kk].abs().max()) < 1e-3
            continue
        cmp(a["params"][kk], b["params"][kk], "d " + kk)


def test_hot_path_net_train_step_native_vs_decomposition():
    """The whole trainable hot path (ptt_b200/train.py::HotPathNet): one forward + backward on the native path against
    the torch decomposition -- loss, a sample of gradients, BatchNorm running statistics (updated TWICE per step by the
    siamese backbone, pointnet2_backbone.py:56-62)."""
    from ptt_b200 import train
    search = torch.from_numpy(synth.make_clouds(4, 1024, 800, "dense")).to(DEV)
    template = torch.from_numpy(synth.make_clouds(4, 512, 801, "dense", role="template")).to(DEV)
    got = []
    for native in (True, False):
        net = train.HotPathNet()
        synth.load_filled(net, seed=0)
        net = net.to(DEV).train()
        for m in net.modules():
            if hasattr(m, "native_train"):
                m.native_train = native
        out = net(search, template)
        loss = sum((v.float() ** 2).mean() for v in out.values())
        loss.backward()
        got.append((float(loss), {k: p.grad for k, p in net.named_parameters()}, {k: b.clone() for k, b in net.named_buffers()}))
    (la, ga, ba), (lb, gb, bb) = got
    assert abs(la - lb) <= 1e-4 * abs(lb), (la, lb)
    assert set(ga) == set(gb) and all(g is not None for g in ga.values())
    for k in gb:
        if k.endswith("fc_gamma.2.bias"):
            continue
        close_grad(ga[k], gb[k], "d " + k, cos_tol=1e-3, l2_tol=5e-2)
    for k in bb:
        close(ba[k], bb[k], 1e-4, k)


@pytest.mark.parametrize("R,d", [(40000, 512), (5000, 256), (12345, 512), (300, 512), (2000, 128)])
def test_rows_linear_masked_vs_torch(R, d):
    """ptt_tr_rows_linear (persistent CTA-pair kernel, plain rows in, ReLU-backward mask in the epilogue) and its fallback
    (d = 128: generic contraction + ptt_tr_mask_positive) against float64."""
    rs = np.random.RandomState(R + d)
    x = torch.from_numpy(rs.standard_normal((R, d)).astype(np.float32)).to(DEV)
    w = torch.from_numpy((rs.standard_normal((d, d)) / np.sqrt(d)).astype(np.float32)).to(DEV)
    ref = torch.from_numpy(rs.standard_normal((R, d)).astype(np.float32)).to(DEV)
    ref[::7] = 0.0                                            # exact zeros mask the gradient (relu'(0) = 0)
    packed = ops.PackedLinear(w.t(), None)                    # out = x . W  (the input gradient of y = h . W^T)
    want = x.double() @ w.double()
    close(train_ops.rows_linear_masked(packed, x), want, 2e-5, "plain")
    close(train_ops.rows_linear_masked(packed, x, ref), want * (ref > 0), 2e-5, "masked")


def test_batched_repack_equals_per_layer_packs():
    """train_ops.repack_stale: after an in-place parameter update every cached image (forward and transposed) is refreshed
    by ONE ptt_linear_pack_batch launch and equals a fresh per-layer pack."""
    from ptt_b200 import train
    net = train.HotPathNet()
    synth.load_filled(net, seed=0)
    net = net.to(DEV).train()
    search = torch.from_numpy(synth.make_clouds(2, 1024, 810, "dense")).to(DEV)
    template = torch.from_numpy(synth.make_clouds(2, 512, 811, "dense", role="template")).to(DEV)
    sum((v.float() ** 2).mean() for v in net(search, template).values()).backward()      # creates the caches
    mods = [m for m in net.modules() if getattr(m, "train_packs", None) is not None]
    n_entries = sum(len(m.train_packs.entries) for m in mods)
    assert len(mods) >= 5 and n_entries >= 40
    assert train_ops.repack_stale(net.modules()) == 0, "nothing is stale right after the step"
    with torch.no_grad():
        for p in net.parameters():
            p.mul_(1.25).add_(0.01)                                                      # what an optimiser step does
    assert train_ops.repack_stale(net.modules()) == n_entries
    for m in mods:
        for ent in m.train_packs.entries.values():
            w2 = ent["weight"].detach().reshape(ent["shape"])
            fresh = ops.PackedLinear(w2.t() if ent["t"] else w2, None if ent["bias"] is None else ent["bias"].detach(), check_range=False)
            assert torch.equal(ent["packed"].params, fresh.params)
    train_ops.invalidate_packs(net.modules())                 # what a caller does after CUDA-graph replays of the step
    assert train_ops.repack_stale(net.modules()) == n_entries


def test_train_step_cuda_graph_replay_matches_eager_steps():
    """train.time_train_step: the whole step (forward, backward, clipping, Adam) captured once and replayed must follow the
    same loss trajectory as eager launches of the same step (the split-K atomics are the only unordered arithmetic)."""
    from ptt_b200 import train
    runs = {}
    for graph in (True, False):
        r = train.time_train_step(torch.device(DEV), batch=4, steps=4, warmup=3, graph=graph)
        runs[graph] = r
        assert ("CUDA graph" in r["launch_mode"]) == graph, r["launch_mode"]
    a, b = runs[True]["loss_first_last"], runs[False]["loss_first_last"]
    assert all(np.isfinite(a)) and a[1] < a[0], "the loss goes down"
    # two runs of the same steps differ through the unordered split-K atomics and the ReLU masks they flip; the trajectories
    # stay together to a few 1e-3 over these 9 updates (a broken capture -- a missing kernel, a stale input -- is off by far more)
    assert abs(a[0] - b[0]) <= 1e-2 * abs(b[0]) and abs(a[1] - b[1]) <= 3e-2 * abs(b[1]), (a, b)
    assert runs[True]["host_enqueue_ms_per_step"] < runs[True]["ms_per_step"], "a replay is enqueued faster than it runs"


@pytest.mark.parametrize("R,K,N", [(50000, 128, 128), (40001, 64, 64), (33000, 3, 64), (20000, 131, 128), (12345, 260, 128),
                                   (30000, 128, 256), (9000, 256, 131), (5000, 64, 3), (70000, 128, 64), (2000, 128, 128),
                                   (8000, 256, 512)])
def test_linear_with_stats_weight_stationary_and_fallback(R, K, N):
    """ptt_linear_fwd_stats: the weight-stationary persistent kernel (bias-free, >= 4096 rows, N <= 256) and the tc_gemm +
    column-reduction fallback (small R, N = 512, biased layers) against torch, incl. the fused statistics, the operand
    transform, ragged last tiles and padded output rows."""
    rs = np.random.RandomState(R % 1000 + K)
    ldx, ldy = (K + 3) // 4 * 4, (N + 3) // 4 * 4
    x = torch.from_numpy(rs.standard_normal((R, ldx)).astype(np.float32)).to(DEV)
    w = torch.from_numpy((rs.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)).to(DEV)
    ka = torch.from_numpy(rs.uniform(0.5, 1.5, K).astype(np.float32)).to(DEV)
    kb = torch.from_numpy(rs.normal(0, 0.5, K).astype(np.float32)).to(DEV)
    for bias in (None, torch.from_numpy(rs.normal(0, 0.3, N).astype(np.float32)).to(DEV)):
        lin = ops.PackedLinear(w, bias)
        for aff in (None, (ka, kb)):
            xin = x[:, :K].double() if aff is None else torch.relu(x[:, :K] * ka + kb).double()
            want = xin @ w.double().t() + (bias.double() if bias is not None else 0.0)
            ws = bias is None and R >= 4096 and N <= 256 and -(-K // 64) * (2 if N > 128 else 1) * 32768 + 65536 + 1280 <= 230400
            stats = ws or N % 4 == 0                    # the fallback's column reduction works in float4 granules
            y, sums = ops.linear_with_stats(lin, x, in_affine=aff, ld_out=ldy, want_stats=stats)
            close(y[:, :N], want, 2e-5, "y")
            assert ldy == N or not y[:, N:].any()
            if stats:
                close(sums[0], want.sum(0), 1e-4, "sum y")
                close(sums[1], (want * want).sum(0), 1e-4, "sum y^2")
