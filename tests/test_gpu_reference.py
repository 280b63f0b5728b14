"""GPU box, with the reference tree staged by oracle/make_ref.sh (oracle/_ref travels with the snapshot): the
REFERENCE'S OWN code on the B200, over the product drop-in.

  Level 1   the reference's own wrappers (pointnet2_utils.QueryAndGroup, PointnetSAModuleVotes, TransformerBlock:
            torch / cuDNN / cuBLAS arithmetic) call `pointnet2_ops._ext` = ptt_b200/dropin -> C ABI -> our kernels;
            results against the CPU oracle (bit-exact where only copies / IEEE elementwise ops are involved) and
            against the fused modules.
  Level 2   the reference's own `build_network` (ptt/models/__init__.py:9-10) with ptt_b200.modules.register():
            the tracker it assembles contains the fused modules and its own PTT.forward (trackers/ptt.py:42-51)
            drives them; `pred_box_data` against HotPath.forward_full and against the reference's own classes.

cuDNN / cuBLAS run the reference's convolutions and linears in TF32 by default; the comparisons switch that off
(the fused path is fp32-class either way)."""
import numpy as np
import pytest
import torch

from conftest import t
from oracle import cops, ref_hotpath, refload, torch_port
from ptt_b200 import hotpath, modules, synth

pytestmark = [pytest.mark.gpu, pytest.mark.reference]

DEV = "cuda:0"
FP_TOL = dict(rtol=1e-4, atol=1e-4)


def g(a):
    if isinstance(a, np.ndarray):
        a = t(a)
    return a.to(DEV).contiguous()


@pytest.fixture(autouse=True)
def _true_fp32():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


@pytest.fixture(scope="module")
def ref_models():
    return refload.load(DEV)


def test_reference_wrappers_resolve_to_the_dropin(ref_models):
    import ptt.models.backbones_3d.pointnet2.pointnet2_utils as pu
    import ptt_b200
    assert pu._ext.__file__.startswith(ptt_b200.DROPIN_DIR)
    assert refload.REFERENCE_ROOT.endswith(("oracle/_ref", "reference"))


def test_reference_query_and_group_over_dropin_is_bit_exact(ref_models):
    """pointnet2_utils.QueryAndGroup.forward (:320-380) on the GPU: ball_query + 2 x group_points are ours, the
    centre subtraction / cat are torch's -- copies and one correctly rounded fp32 subtraction, so without
    NORMALIZE_XYZ the result equals the CPU oracle's bit for bit.  With it, torch's CUDA kernel divides by a Python
    scalar as x * (1 / radius) (<= 1 ulp from the IEEE quotient the CPU path and our fused kernel produce)."""
    import ptt.models.backbones_3d.pointnet2.pointnet2_utils as pu
    for seed, kind, n, m, c, radius, ns in ((1, "dense", 1024, 512, 0, 0.3, 32), (2, "sparse", 512, 256, 16, 0.5, 32),
                                            (3, "dense", 128, 64, 257, 0.3, 16)):
        xyz = synth.make_clouds(3, n, 40 + seed, kind)
        feats = synth.features((3, c, n), seed=50 + seed) if c else None
        idx = cops.furthest_point_sampling(t(xyz), m).long()
        new_xyz = np.take_along_axis(xyz, idx.numpy()[:, :, None], 1)
        for normalize in (False, True):
            qg = pu.QueryAndGroup(radius, ns, use_xyz=True, ret_grouped_xyz=True, normalize_xyz=normalize)
            got, got_xyz = qg(g(xyz), g(new_xyz), g(feats) if c else None)
            want, want_xyz, _ = torch_port.query_and_group(t(xyz), t(new_xyz), t(feats) if c else None, radius, ns, True,
                                                           normalize)
            if normalize:
                np.testing.assert_allclose(got_xyz.cpu().numpy(), want_xyz.numpy(), rtol=3e-7, atol=0)
                assert np.array_equal(got[:, 3:].cpu().numpy(), want[:, 3:].numpy())
            else:
                assert np.array_equal(got.cpu().numpy(), want.numpy())
                assert np.array_equal(got_xyz.cpu().numpy(), want_xyz.numpy())


def test_reference_sa_module_over_dropin_matches_fused_module(ref_models, golden):
    """The reference's own PointnetSAModuleVotes (cuDNN 1x1 convolutions + BatchNorm + max_pool over our `_ext`) and the
    fused ptt_b200 module with the same state_dict: identical samples, features within 1e-4 of each other and of the
    committed fixture the reference produced on CPU."""
    from ptt.models.backbones_3d.pointnet2.pointnet2_modules import PointnetSAModuleVotes as RefSA
    from test_oracle_golden import SA_CASES, sa_state_dict
    gd = golden("sa_module.npz")
    for i, (name, (n, cin, mlp, npoint, radius, ns, method)) in enumerate(SA_CASES.items()):
        sd = {k: t(v) for k, v in synth.fill_state_dict(sa_state_dict(mlp), seed=10 + i).items()}
        kw = dict(radius=radius, nsample=ns, use_xyz=True, normalize_xyz=True, sample_method=method)
        ref = RefSA(mlp=list(mlp), **kw)
        ours = modules.PointnetSAModuleVotes(mlp=list(mlp), **kw)
        ref.load_state_dict(sd), ours.load_state_dict(sd)
        ref, ours = ref.to(DEV).eval(), ours.to(DEV).eval()
        xyz = g(gd[name + "/xyz"])
        feats = g(synth.features((2, cin, n), seed=30 + i)) if cin else None
        with torch.no_grad():
            r_xyz, r_f, r_i = ref(xyz, feats, npoint)
            o_xyz, o_f, o_i = ours(xyz, feats, npoint)
        assert torch.equal(r_i, o_i) and torch.equal(r_xyz, o_xyz), name
        assert np.array_equal(r_i.cpu().numpy(), gd[name + "/inds"]), name
        np.testing.assert_allclose(r_f.cpu().numpy(), gd[name + "/new_features"], err_msg=name, **FP_TOL)
        np.testing.assert_allclose(o_f.cpu().numpy(), r_f.cpu().numpy(), err_msg=name, **FP_TOL)


def test_reference_transformer_block_on_gpu_matches_fused_block(ref_models):
    from ptt.models import transformer_block
    from test_oracle_golden import transformer_state_dict
    for name, n, dp, dm, k in (("TransformerBlock", 128, 256, 512, 16), ("TransformerBlockOffset", 64, 64, 128, 8)):
        sd = {kk: t(v) for kk, v in synth.fill_state_dict(transformer_state_dict(name, dp, dm), seed=77).items()}
        ref = transformer_block.__all__[name](d_points=dp, d_model=dm, k=k)
        assert type(ref).__module__.startswith("ptt.models")          # the reference's class, not ours
        ours = modules.REGISTRY[name](dp, dm, k)
        ref.load_state_dict(sd), ours.load_state_dict(sd)
        ref, ours = ref.to(DEV).eval(), ours.to(DEV).eval()
        xyz = g(synth.make_clouds(3, n, 78, "dense", role="template"))      # no exact ties: argsort's order is defined
        f = g(synth.features((3, n, dp), seed=79))
        with torch.no_grad():
            r_out, r_attn = ref(xyz, f)
            o_out, o_attn = ours(xyz, f)
        np.testing.assert_allclose(o_out.cpu().numpy(), r_out.cpu().numpy(), err_msg=name, **FP_TOL)
        np.testing.assert_allclose(o_attn.cpu().numpy(), r_attn.cpu().numpy(), err_msg=name, **FP_TOL)


def test_reference_hot_path_on_gpu_matches_hotpath(ref_models):
    """The whole hot path through the reference's own module objects on the B200 (over the drop-in) vs HotPath."""
    sd = synth.full_model_state_dict(0)
    ref = ref_hotpath.RefHotPath(sd, DEV)
    hp = hotpath.HotPath(sd, device=DEV)
    for kind, ns, nt, seed in (("dense", 1024, 512, 820), ("sparse", 512, 512, 830)):
        search, template = g(synth.make_clouds(4, ns, seed, kind)), g(synth.make_clouds(4, nt, seed + 1, kind, role="template"))
        want = ref.hot_path(search, template)
        got = hp(search, template)
        torch.cuda.synchronize()
        for k in ("search_inds", "template_inds"):
            assert torch.equal(got[k], want[k]), (kind, k)
        for k in ("search_seeds", "template_seeds", "box_centers"):
            assert torch.equal(got[k], want[k]), (kind, k)
        for k in ("search_feats", "template_feats", "centroid_feats", "box_sa_feats", "box_feats"):
            np.testing.assert_allclose(got[k].cpu().numpy(), want[k].cpu().numpy(), err_msg="%s %s" % (kind, k), **FP_TOL)


def test_reference_build_network_with_registered_b200_modules(ref_models):
    """Level 2: the reference's own factory + its own PTT.forward, with the fused modules registered, on the B200."""
    sd = synth.full_model_state_dict(0)
    reg = ref_hotpath.RefHotPath(sd, DEV, register_b200=True)
    net = reg.net
    assert isinstance(net.backbone_3d.SA_modules[0], modules.PointnetSAModuleVotes)
    assert isinstance(net.centroid_voting_head.transformer_block, modules.TransformerBlock)
    assert isinstance(net.box_voting_head.vote_aggregation, modules.PointnetSAModuleVotes)
    assert isinstance(net.box_voting_head.transformer_block, modules.TransformerBlock)
    own = ref_hotpath.RefHotPath(sd, DEV)                        # the reference's own classes (over the `_ext` drop-in)
    assert type(own.net.backbone_3d.SA_modules[0]).__module__.startswith("ptt.models")
    hp = hotpath.HotPath(sd, device=DEV)
    search, template = g(synth.make_clouds(4, 1024, 840, "dense")), g(synth.make_clouds(4, 512, 841, "dense", role="template"))
    a = reg.full(search, template)                               # reference forward, fused modules inside
    b = own.full(search, template)                               # reference forward, reference modules
    c = hp.forward_full(search, template)                        # the product's whole-tracker forward
    torch.cuda.synchronize()
    for k in ("search_feats", "template_feats", "cosine_feats", "pred_centroids_cls", "pred_centroids_votes"):
        np.testing.assert_allclose(a[k].cpu().numpy(), b[k].cpu().numpy(), err_msg=k, **FP_TOL)
        np.testing.assert_allclose(c[k].cpu().numpy(), b[k].cpu().numpy(), err_msg=k, **FP_TOL)
    # the box head starts with FPS over PREDICTED votes (a 1e-6 difference may flip a pick): compare the final boxes only
    # where the three runs picked the same centres -- on these inputs they do
    for x in (a, c):
        assert torch.allclose(x["pred_box_center"], b["pred_box_center"], rtol=1e-5, atol=1e-4)
        np.testing.assert_allclose(x["pred_box_data"].cpu().numpy(), b["pred_box_data"].cpu().numpy(), **FP_TOL)
    est = b["pred_box_data"]
    pick = est[:, :, 4].argmax(1)
    assert torch.equal(c["best_idx"], pick)
