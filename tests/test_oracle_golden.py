"""CPU: the oracle (C ops + torch port) against the committed golden fixtures the reference produced.

ops.npz is oracle-generated (rows a1-a4 are parity-unpinned: upstream pointnet2_ops is unavailable) and
guards against drift; sa_module / transformer / hot_path fixtures were produced by the REFERENCE's own
modules (tests/golden/make_golden.py) and pin oracle/torch_port.py.
"""
import numpy as np
import pytest
import torch

from conftest import t
from oracle import cops, torch_port
from ptt_b200 import synth

FP_TOL = dict(rtol=1e-4, atol=1e-4)   # north_star: fp32 features within 1e-4


def test_ops_fixture_matches_c_oracle(golden):
    g = golden("ops.npz")
    names = sorted({k.split("/")[0] for k in g.files})
    assert len(names) == 6
    for name in names:
        xyz = t(g[name + "/xyz"])
        m = xyz.shape[1] // 2
        idx = cops.furthest_point_sampling(xyz, m)
        assert np.array_equal(idx.numpy(), g[name + "/fps"]), name
        new_xyz = cops.gather_points(xyz.transpose(1, 2).contiguous(), idx).transpose(1, 2).contiguous()
        for r, ns in ((0.3, 32), (0.7, 16)):
            assert np.array_equal(cops.ball_query(new_xyz, xyz, r, ns).numpy(), g[name + "/bq_r%g_ns%d" % (r, ns)])
        assert np.array_equal(cops.knn(new_xyz, min(16, m)).numpy(), g[name + "/knn16"])
        d2, i3 = cops.three_nn(xyz, new_xyz)
        assert np.array_equal(i3.numpy(), g[name + "/three_nn_idx"])
        assert np.array_equal(d2.numpy(), g[name + "/three_nn_d2"])


def test_fps_known_answers():
    # hand-checkable cases of the published algorithm
    # (1) four collinear points: start at 0, then the farthest, then the one maximising the min distance
    xyz = torch.tensor([[[1.0, 0, 0], [2.0, 0, 0], [4.0, 0, 0], [8.0, 0, 0]]])
    assert cops.furthest_point_sampling(xyz, 4).tolist() == [[0, 3, 2, 1]]
    # (2) all-zero cloud: every point is inside the origin ball and skipped -> always index 0
    assert cops.furthest_point_sampling(torch.zeros(1, 16, 3), 5).tolist() == [[0] * 5]
    # (3) exact tie between slots 1 and 2 of a 4-thread block: the tree keeps slot 2 (bit-reversed order)
    xyz = torch.tensor([[[1.0, 0, 0], [1.0, 3.0, 0], [1.0, -3.0, 0], [1.0, 0.5, 0]]])
    assert cops.furthest_point_sampling(xyz, 2).tolist() == [[0, 2]]
    # (4) a point inside the 1e-3 origin ball is never selected while others remain
    xyz = torch.tensor([[[1.0, 0, 0], [0.01, 0.0, 0.0], [2.0, 0, 0], [3.0, 0, 0]]])
    assert 1 not in cops.furthest_point_sampling(xyz, 3).tolist()[0]


def test_ball_query_known_answers():
    xyz = torch.tensor([[[0.0, 0, 0], [0.1, 0, 0], [0.2, 0, 0], [5.0, 0, 0], [0.05, 0, 0]]])
    centres = torch.tensor([[[0.0, 0, 0], [9.0, 9.0, 9.0], [5.0, 0, 0]]])
    idx = cops.ball_query(centres, xyz, 0.15, 4)
    # in-order first hits, the first hit pads the tail; no hit -> zeros
    assert idx.tolist() == [[[0, 1, 4, 0], [0, 0, 0, 0], [3, 3, 3, 3]]]
    # strict '<': a point exactly on the radius is outside
    idx = cops.ball_query(torch.zeros(1, 1, 3), torch.tensor([[[0.5, 0, 0], [0.25, 0, 0]]]), 0.5, 2)
    assert idx.tolist() == [[[1, 1]]]


def test_group_and_gather_grad_are_adjoint():
    rs = np.random.RandomState(0)
    pts = t(rs.standard_normal((2, 5, 17)).astype(np.float32))
    idx = t(rs.randint(0, 17, size=(2, 6, 4)).astype(np.int32))
    g = t(rs.standard_normal((2, 5, 6, 4)).astype(np.float32))
    lhs = (cops.group_points(pts, idx) * g).sum()
    rhs = (pts * cops.group_points_grad(g, idx, 17)).sum()
    assert torch.allclose(lhs, rhs, rtol=1e-5)
    idx2 = t(rs.randint(0, 17, size=(2, 9)).astype(np.int32))
    g2 = t(rs.standard_normal((2, 5, 9)).astype(np.float32))
    assert torch.allclose((cops.gather_points(pts, idx2) * g2).sum(), (pts * cops.gather_points_grad(g2, idx2, 17)).sum(), rtol=1e-5)


SA_CASES = {
    "sa1": (1024, 0, [0, 64, 64, 128], 512, 0.3, 32, "fps"),
    "sa2": (512, 128, [128, 128, 128, 256], 256, 0.5, 32, "sequence"),
    "sa3_sparse": (256, 256, [256, 128, 128, 256], 128, 0.7, 32, "sequence"),
    "box": (128, 257, [257, 256, 256, 256], 64, 0.3, 16, "fps"),
    "ragged": (200, 5, [5, 24, 40], 50, 0.4, 8, "fps"),
}


def sa_state_dict(mlp, use_xyz=True):
    """Key/shape layout of PointnetSAModuleVotes.state_dict()."""
    spec = list(mlp)
    if use_xyz:
        spec[0] += 3
    sd = {}
    for i in range(len(spec) - 1):
        p = "mlp_module.layer%d." % i
        sd[p + "conv.weight"] = torch.empty(spec[i + 1], spec[i], 1, 1)
        for k in ("weight", "bias", "running_mean", "running_var"):
            sd[p + "normlayer.bn." + k] = torch.empty(spec[i + 1])
        sd[p + "normlayer.bn.num_batches_tracked"] = torch.empty((), dtype=torch.int64)
    return sd


@pytest.mark.parametrize("name", list(SA_CASES))
def test_port_sa_module_vs_reference_fixture(golden, name):
    g = golden("sa_module.npz")
    i = list(SA_CASES).index(name)
    n, cin, mlp, npoint, radius, ns, method = SA_CASES[name]
    sd = {k: t(v) for k, v in synth.fill_state_dict(sa_state_dict(mlp), seed=10 + i).items()}
    xyz = t(g[name + "/xyz"])
    feats = None
    if cin:
        f = synth.features((2, cin, n), seed=30 + i)
        assert synth.crc(f) == int(g[name + "/features_crc"])
        feats = t(f)
    new_xyz, new_feats, inds = torch_port.sa_module_votes(sd, xyz, feats, npoint, radius, ns, method, True, True)
    assert np.array_equal(inds.numpy(), g[name + "/inds"])
    assert np.array_equal(new_xyz.numpy(), g[name + "/new_xyz"])
    np.testing.assert_allclose(new_feats.numpy(), g[name + "/new_features"], **FP_TOL)


TR_CASES = {
    "centroid": ("TransformerBlock", 128, 256, 512, 16),
    "box": ("TransformerBlock", 64, 256, 512, 16),
    "small": ("TransformerBlock", 40, 24, 48, 5),
    "mlp": ("TransformerBlockMLP", 32, 32, 64, 8),
    "offset": ("TransformerBlockOffset", 32, 32, 64, 8),
    "std": ("TransformerBlockSTD", 48, 32, 64, 8),
}


def transformer_state_dict(cls, dp, dm):
    sd = {}

    def lin(name, i, o, bias=True):
        sd[name + ".weight"] = torch.empty(o, i)
        if bias:
            sd[name + ".bias"] = torch.empty(o)

    if cls == "TransformerBlockMLP":
        lin("fc1.0", dp, dm), lin("fc1.2", dm, dm), lin("fc2.0", dm, dm), lin("fc2.2", dm, dp)
    else:
        lin("fc1", dp, dm), lin("fc2", dm, dp)
    lin("fc_delta.0", 3, dm), lin("fc_delta.2", dm, dm)
    if cls != "TransformerBlockSTD":
        lin("fc_gamma.0", dm, dm), lin("fc_gamma.2", dm, dm)
    lin("w_qs", dm, dm, False), lin("w_ks", dm, dm, False), lin("w_vs", dm, dm, False)
    return sd


@pytest.mark.parametrize("name", list(TR_CASES))
def test_port_transformer_vs_reference_fixture(golden, name):
    g = golden("transformer.npz")
    i = list(TR_CASES).index(name)
    cls, n, dp, dm, k = TR_CASES[name]
    sd = {kk: t(v) for kk, v in synth.fill_state_dict(transformer_state_dict(cls, dp, dm), seed=40 + i).items()}
    xyz = t(g[name + "/xyz"])
    f = synth.features((2, n, dp), seed=60 + i)
    assert synth.crc(f) == int(g[name + "/features_crc"])
    if cls == "TransformerBlockSTD":
        res, attn = torch_port.transformer_block_std(sd, xyz, t(f))
    else:
        res, attn = torch_port.transformer_block(sd, xyz, t(f), k, variant=cls)
    np.testing.assert_allclose(res.numpy(), g[name + "/res"], **FP_TOL)
    np.testing.assert_allclose(attn[:, :4].numpy(), g[name + "/attn_head"], **FP_TOL)


FULL_KEYS = ("search_feats", "template_feats", "cosine_feats", "pred_centroids_cls", "pred_centroids_votes",
             "pred_box_center", "pred_box_data")


def test_port_full_model_vs_reference_fixture(golden):
    """The whole tracker forward of the port (backbone -> CosineSimAug -> both heads) against
    the outputs of the REFERENCE's own PTT model (hot_path.npz full/*, made by make_golden.py)."""
    g = golden("hot_path.npz")
    sd = synth.full_model_state_dict(0)
    out = torch_port.full_model_frame(sd, t(g["full/search"]), t(g["full/template"]))
    # FPS over the predicted votes: same picks (the values themselves carry the votes' 1e-6 summation-order noise)
    np.testing.assert_allclose(out["pred_box_center"].numpy(), g["full/pred_box_center"], rtol=1e-5, atol=1e-4)
    for k in FULL_KEYS:
        np.testing.assert_allclose(out[k].numpy(), g["full/" + k], err_msg=k, **FP_TOL)


# the other registered blocks: name -> (class, n, d_points, d_model, k, heads, layers); fixtures made by make_golden.py
EXTRA_TR_CASES = {
    "cosine": ("TransformerBlockCosine", 48, 32, 64, 8, 1, 1),
    "all": ("TransformerBlockALL", 40, 24, 48, 4, 1, 1),
    "cross": ("CrossAttentionBlock", 48, 32, 64, 8, 1, 1),
    "mul": ("MulTransformerBlock", 48, 32, 64, 8, 4, 2),
}


def extra_transformer_state_dict(cls, dp, dm, heads=1, layers=1):
    """state_dict layout (key -> shape) of the secondary blocks (variants.py:43-124,168-208; multitransformer.py:11-76)."""
    if cls == "MulTransformerBlock":
        hd = dm // heads
        one = transformer_state_dict("TransformerBlock", dp, dm)
        for nm in ("fc_gamma.0", "fc_gamma.2"):
            one[nm + ".weight"], one[nm + ".bias"] = (hd, hd), (hd,)
        one.update({"proj.weight": (dm, dm), "norm1.weight": (dm,), "norm1.bias": (dm,), "norm2.weight": (dp,), "norm2.bias": (dp,)})
        return {"layers.%d.%s" % (i, k): v for i in range(layers) for k, v in one.items()}
    sd = transformer_state_dict("TransformerBlock", dp, dm)
    if cls == "TransformerBlockCosine":
        sd.update({"fc_sim.weight": (dm, dm + 1), "fc_sim.bias": (dm,)})
    if cls == "CrossAttentionBlock":
        sd.update({"fc2.weight": (dm, dp), "fc2.bias": (dm,), "fc3.weight": (dp, dm), "fc3.bias": (dp,)})
    return sd


def run_extra_port(cls, sd, xyz, feats, k, heads, feats2=None):
    if cls == "TransformerBlockCosine":
        return torch_port.transformer_block_cosine(sd, xyz, feats, k)
    if cls == "TransformerBlockALL":
        return torch_port.transformer_block_all(sd, xyz, feats)
    if cls == "CrossAttentionBlock":
        return torch_port.cross_attention_block(sd, xyz, feats, feats2, k)
    return torch_port.mul_transformer_block(sd, xyz, feats, k, heads)


@pytest.mark.parametrize("name", list(EXTRA_TR_CASES))
def test_port_secondary_blocks_vs_reference_fixture(golden, name):
    g = golden("transformer.npz")
    i = list(EXTRA_TR_CASES).index(name)
    cls, n, dp, dm, k, heads, layers = EXTRA_TR_CASES[name]
    sd = {kk: t(v) for kk, v in synth.fill_state_dict(extra_transformer_state_dict(cls, dp, dm, heads, layers), seed=80 + i).items()}
    f = synth.features((2, n, dp), seed=100 + i)
    assert synth.crc(f) == int(g[name + "/features_crc"])
    f2 = t(synth.features((2, n, dp), seed=110 + i)) if cls == "CrossAttentionBlock" else None
    res, attn = run_extra_port(cls, sd, t(g[name + "/xyz"]), t(f), k, heads, f2)
    np.testing.assert_allclose(res.numpy(), g[name + "/res"], **FP_TOL)
    np.testing.assert_allclose(attn[:, :4].numpy(), g[name + "/attn_head"], **FP_TOL)
