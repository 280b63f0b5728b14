"""GPU parity: the sm_100a kernels behind the `pointnet2_ops._ext` surface vs the CPU oracle.

Everything goes through the C ABI (ptt_b200.ops -> ctypes -> libptt_b200.so).  Integer / index work is
compared bit for bit; gradients (atomic accumulation order differs) within 1e-5.
"""
import numpy as np
import pytest
import torch

from conftest import t
from oracle import cops
from ptt_b200 import ops, synth

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def g(a):
    if isinstance(a, np.ndarray):
        a = t(a)
    return a.to(DEV).contiguous()


def test_library_reports_sm100a():
    from ptt_b200 import _lib
    assert "sm_100a" in _lib.version()
    assert torch.cuda.get_device_capability(0)[0] == 10


def test_ops_fixture_bit_exact(golden):
    gd = golden("ops.npz")
    names = sorted({k.split("/")[0] for k in gd.files})
    for name in names:
        xyz = g(gd[name + "/xyz"])
        m = xyz.shape[1] // 2
        idx, new_xyz = ops.furthest_point_sampling(xyz, m, return_new_xyz=True)
        assert np.array_equal(idx.cpu().numpy(), gd[name + "/fps"]), name
        want_new = np.take_along_axis(gd[name + "/xyz"], gd[name + "/fps"].astype(np.int64)[:, :, None], 1)
        assert np.array_equal(new_xyz.cpu().numpy(), want_new), name
        for r, ns in ((0.3, 32), (0.7, 16)):
            got = ops.ball_query(new_xyz, xyz, r, ns).cpu().numpy()
            assert np.array_equal(got, gd[name + "/bq_r%g_ns%d" % (r, ns)]), (name, r, ns)
        assert np.array_equal(ops.knn(new_xyz, min(16, m)).cpu().numpy(), gd[name + "/knn16"]), name
        d2, i3 = ops.three_nn(xyz, new_xyz)
        assert np.array_equal(i3.cpu().numpy(), gd[name + "/three_nn_idx"]), name
        assert np.array_equal(d2.cpu().numpy(), gd[name + "/three_nn_d2"]), name


@pytest.mark.parametrize("n,m,kind", [(1024, 512, "dense"), (512, 512, "sparse"), (2048, 1024, "dense"),
                                      (256, 128, "sparse"), (128, 64, "dense"), (1000, 333, "dense"),
                                      (37, 37, "sparse"), (5000, 100, "dense"), (9000, 64, "dense"), (1, 1, "dense")])
def test_fps_vs_oracle(n, m, kind):
    xyz = synth.make_clouds(4, n, 100 + n, kind)
    want = cops.furthest_point_sampling(t(xyz), m).numpy()
    got, new_xyz = ops.furthest_point_sampling(g(xyz), m, return_new_xyz=True)
    assert np.array_equal(got.cpu().numpy(), want)
    assert np.array_equal(new_xyz.cpu().numpy(), np.take_along_axis(xyz, want.astype(np.int64)[:, :, None], 1))


def test_fps_adversarial_ties():
    for n in (64, 256, 1024):
        xyz = synth.adversarial_clouds(n, seed=n)
        want = cops.furthest_point_sampling(t(xyz), n // 2).numpy()
        got = ops.furthest_point_sampling(g(xyz), n // 2).cpu().numpy()
        assert np.array_equal(got, want), n


def test_fps_every_tuning_variant_agrees():
    import ctypes
    from ptt_b200 import _lib
    L = _lib.lib()
    fn = L.ptt_fps_variant
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                   ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    xyz = np.concatenate([synth.make_clouds(2, 512, 5, "sparse"), synth.adversarial_clouds(512, 1)], 0)
    want = cops.furthest_point_sampling(t(xyz), 256).numpy()
    dx = g(xyz)
    for threads in (32, 64, 128, 256, 512):
        for ppt in (1, 2, 4, 8, 16):
            if threads * ppt < 512:
                continue
            out = torch.full((xyz.shape[0], 256), -1, dtype=torch.int32, device=DEV)
            rc = fn(dx.data_ptr(), xyz.shape[0], 512, 256, out.data_ptr(), None, threads, ppt,
                    torch.cuda.current_stream().cuda_stream)
            if rc == -2:
                continue
            assert rc == 0
            assert np.array_equal(out.cpu().numpy(), want), (threads, ppt)


def test_fps_with_dist_vs_oracle():
    rs = np.random.RandomState(3)
    p = rs.standard_normal((3, 96, 5)).astype(np.float32)
    d = ((p[:, :, None] - p[:, None]) ** 2).sum(-1).astype(np.float32)
    want = cops.furthest_point_sampling_with_dist(t(d), 40).numpy()
    got = ops.furthest_point_sampling_with_dist(g(d), 40).cpu().numpy()
    assert np.array_equal(got, want)


@pytest.mark.parametrize("n,m,r,ns,kind", [(1024, 512, 0.3, 32, "dense"), (512, 256, 0.5, 32, "dense"),
                                           (256, 128, 0.7, 32, "sparse"), (128, 64, 0.3, 16, "dense"),
                                           (1000, 77, 0.4, 5, "dense"), (2048, 1024, 0.3, 64, "sparse"),
                                           (20000, 50, 0.2, 16, "dense"), (33, 33, 10.0, 48, "dense")])
def test_ball_query_vs_oracle(n, m, r, ns, kind):
    xyz = synth.make_clouds(3, n, 200 + n, kind)
    centres = xyz[:, np.random.RandomState(n).permutation(n)[:m]].copy()
    centres[:, -1] = 100.0                        # a centre with no neighbour at all -> all-zero row
    want = cops.ball_query(t(centres), t(xyz), r, ns).numpy()
    got = ops.ball_query(g(centres), g(xyz), r, ns).cpu().numpy()
    assert np.array_equal(got, want)
    assert (got[:, -1] == 0).all()


@pytest.mark.parametrize("n,npoints,kind", [(1024, (512, 256, 128), "dense"), (512, (512, 256, 128), "sparse"),
                                            (300, (200, 200, 7), "dense"), (64, (33,), "dense")])
def test_ball_query_nested_equals_the_single_queries(n, npoints, kind):
    """ptt_ball_query_nested (all levels of a backbone branch in one launch) == one ptt_ball_query per level == oracle."""
    xyz = synth.make_clouds(5, n, 230 + n, kind)
    samples = np.take_along_axis(xyz, cops.furthest_point_sampling(t(xyz), npoints[0]).long().numpy()[:, :, None], 1)
    radii, nss = (0.3, 0.5, 0.7)[:len(npoints)], (32, 16, 5)[:len(npoints)]
    got = ops.ball_query_nested(g(xyz), g(samples), npoints, radii, nss)
    src = xyz
    for l, m in enumerate(npoints):
        ctr = np.ascontiguousarray(samples[:, :m])
        want = cops.ball_query(t(ctr), t(src), radii[l], nss[l]).numpy()
        assert np.array_equal(got[l].cpu().numpy(), want), l
        assert np.array_equal(ops.ball_query(g(ctr), g(src), radii[l], nss[l]).cpu().numpy(), want), l
        src = ctr


def test_ball_query_radius_is_strict():
    xyz = np.float32([[[0.5, 0, 0], [0.25, 0, 0]]])
    got = ops.ball_query(g(np.zeros((1, 1, 3), np.float32)), g(xyz), 0.5, 2).cpu().numpy()
    assert got.tolist() == [[[1, 1]]]


def test_gather_group_and_grads_vs_oracle():
    rs = np.random.RandomState(0)
    for (B, C, N, M, K) in ((2, 5, 17, 6, 4), (3, 131, 512, 256, 32), (2, 257, 128, 64, 16), (1, 3, 1024, 512, 32)):
        pts = rs.standard_normal((B, C, N)).astype(np.float32)
        idx2 = rs.randint(0, N, size=(B, M)).astype(np.int32)
        idx3 = rs.randint(0, N, size=(B, M, K)).astype(np.int32)
        assert np.array_equal(ops.gather_points(g(pts), g(idx2)).cpu().numpy(), cops.gather_points(t(pts), t(idx2)).numpy())
        assert np.array_equal(ops.group_points(g(pts), g(idx3)).cpu().numpy(), cops.group_points(t(pts), t(idx3)).numpy())
        g2 = rs.standard_normal((B, C, M)).astype(np.float32)
        g3 = rs.standard_normal((B, C, M, K)).astype(np.float32)
        np.testing.assert_allclose(ops.gather_points_grad(g(g2), g(idx2), N).cpu().numpy(),
                                   cops.gather_points_grad(t(g2), t(idx2), N).numpy(), rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(ops.group_points_grad(g(g3), g(idx3), N).cpu().numpy(),
                                   cops.group_points_grad(t(g3), t(idx3), N).numpy(), rtol=1e-5, atol=2e-5)


def test_group_points_output_is_fresh_memory():
    # the reference mutates the result in place (pointnet2_utils.py:352,354)
    pts = g(np.arange(2 * 3 * 8, dtype=np.float32).reshape(2, 3, 8))
    idx = g(np.zeros((2, 4, 2), np.int32))
    out = ops.group_points(pts, idx)
    before = pts.clone()
    out -= 1.0
    assert torch.equal(pts, before)


def test_three_nn_and_interpolate_vs_oracle():
    rs = np.random.RandomState(1)
    unknown = synth.make_clouds(2, 300, 7, "dense")
    known = synth.make_clouds(2, 1500, 8, "dense")
    d2w, iw = cops.three_nn(t(unknown), t(known))
    d2, i3 = ops.three_nn(g(unknown), g(known))
    assert np.array_equal(i3.cpu().numpy(), iw.numpy()) and np.array_equal(d2.cpu().numpy(), d2w.numpy())
    feats = rs.standard_normal((2, 9, 1500)).astype(np.float32)
    w = rs.uniform(size=(2, 300, 3)).astype(np.float32)
    assert np.array_equal(ops.three_interpolate(g(feats), i3, g(w)).cpu().numpy(),
                          cops.three_interpolate(t(feats), iw, t(w)).numpy())
    go = rs.standard_normal((2, 9, 300)).astype(np.float32)
    np.testing.assert_allclose(ops.three_interpolate_grad(g(go), i3, g(w), 1500).cpu().numpy(),
                               cops.three_interpolate_grad(t(go), iw, t(w), 1500).numpy(), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("n,k", [(128, 16), (64, 16), (40, 5), (300, 16), (1024, 32), (16, 16)])
def test_knn_vs_oracle(n, k):
    xyz = np.concatenate([synth.make_clouds(2, n, 300 + n, "dense", role="template"),
                          synth.make_clouds(2, n, 301 + n, "sparse")], 0)
    want = cops.knn(t(xyz), k).numpy()
    got = ops.knn(g(xyz), k).cpu().numpy()
    assert np.array_equal(got, want)


def test_layout_round_trip():
    x = g(synth.features((3, 37, 101), seed=5))
    pm = ops.cm_to_pm(x, ld=40)
    assert pm.shape == (3, 101, 40)
    assert torch.equal(pm[:, :, :37], x.transpose(1, 2))
    assert (pm[:, :, 37:] == 0).all()
    assert torch.equal(ops.pm_to_cm(pm, 37), x)


def test_linear_vs_torch_fp32():
    rs = np.random.RandomState(2)
    for (R, K, C) in ((1, 3, 5), (100, 259, 128), (4096, 512, 512), (777, 256, 257)):
        x = rs.standard_normal((R, K)).astype(np.float32)
        w = (rs.standard_normal((C, K)) / np.sqrt(K)).astype(np.float32)
        b = rs.standard_normal(C).astype(np.float32)
        res = rs.standard_normal((R, C)).astype(np.float32)
        lin = ops.PackedLinear(g(w), g(b))
        want = torch.relu(t(x).double() @ t(w).double().T + t(b).double()) + t(res).double()
        got = lin(g(x), relu=True, residual=g(res)).cpu().double()
        np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=1e-4, atol=1e-4)   # north_star: fp32 within 1e-4


@pytest.mark.parametrize("R,K,C", [(128, 64, 64), (1, 64, 8), (130, 64, 64), (1000, 260, 200), (4099, 132, 128),
                                   (20000, 512, 512), (333, 512, 1536), (256, 256, 257), (64, 8, 300)])
def test_tensor_core_linear_is_fp32_accurate(R, K, C):
    """The tcgen05 path (fp16 hi/lo split, 3 MMAs per K-step) against float64, and against the CUDA-core path."""
    import ctypes
    from ptt_b200 import _lib
    rs = np.random.RandomState(R + K + C)
    x = (rs.standard_normal((R, K)) * np.exp(rs.uniform(-6, 6, size=(R, 1)))).astype(np.float32)   # rows spanning 5 decades
    w = (rs.standard_normal((C, K)) / np.sqrt(K)).astype(np.float32)
    b = rs.standard_normal(C).astype(np.float32)
    res = rs.standard_normal((R, C)).astype(np.float32)
    lin = ops.PackedLinear(g(w), g(b))
    want = torch.relu(t(x).double() @ t(w).double().T + t(b).double()) + t(res).double()
    got = lin(g(x), relu=True, residual=g(res)).cpu().double()
    scale = (t(x).double().abs() @ t(w).double().abs().T) + 1.0       # conditioning of each dot product
    err = ((got - want).abs() / scale).max().item()
    assert err < 2e-6, err                                            # fp32-class: ~2^-22 relative to sum |a||b|
    force = _lib.lib().ptt_debug_force_ffma
    force.argtypes = [ctypes.c_int]
    force(1)
    try:
        ffma = lin(g(x), relu=True, residual=g(res)).cpu().double()
    finally:
        force(0)
    assert ((ffma - want).abs() / scale).max().item() < 2e-6
    assert ((ffma - got).abs() / scale).max().item() < 2e-6


def test_errors_are_raised_not_fatal():
    xyz = g(synth.make_clouds(1, 64, 1))
    with pytest.raises(RuntimeError):
        ops.furthest_point_sampling(xyz.double(), 8)                  # dtype
    with pytest.raises(RuntimeError):
        ops.furthest_point_sampling(xyz.transpose(1, 2), 8)           # layout / contiguity
    with pytest.raises(RuntimeError):
        ops.ball_query(xyz, xyz.cpu(), 0.3, 4)                        # device
    with pytest.raises(RuntimeError):
        ops.group_points(g(np.zeros((1, 3, 8), np.float32)), g(np.zeros((2, 4, 2), np.int32)))   # batch mismatch
    with pytest.raises(RuntimeError):
        ops.knn(xyz, 65)                                              # k > n
    # empty inputs are fine
    assert ops.furthest_point_sampling(torch.zeros(0, 16, 3, device=DEV), 4).shape == (0, 4)
    assert ops.ball_query(torch.zeros(2, 0, 3, device=DEV), g(synth.make_clouds(2, 16, 0)), 0.3, 4).shape == (2, 0, 4)
