"""Generates tests/golden/*.npz by running the REFERENCE's own modules (imported from /root/reference)
on CPU over the oracle ops.  Run in the authoring container only:

    python tests/golden/make_golden.py

The fixtures pin (a) oracle/torch_port.py and (b) the CUDA path on the GPU box, where the reference
tree does not exist.  Inputs and parameters come from the frozen recipes in ptt_b200/synth.py.
The native-op fixture (ops.npz) is produced by the C oracle itself -- upstream pointnet2_ops is not
available anywhere, so for rows a1-a4 the fixture only guards against oracle drift (parity unpinned).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)

from oracle import cops, refload  # noqa: E402
from ptt_b200 import synth  # noqa: E402

torch.set_grad_enabled(False)
torch.set_num_threads(1)  # fixed summation order for the committed numbers


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def save(name, **arrays):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **{k: (v.numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in arrays.items()})
    print("%-22s %8.1f KiB" % (name, os.path.getsize(path) / 1024))


def ops_fixture():
    out = {}
    clouds = {
        "dense1024": synth.make_clouds(2, 1024, 0, "dense"),
        "sparse1024": synth.make_clouds(3, 1024, 1, "sparse"),
        "sparse512": synth.make_clouds(3, 512, 2, "sparse"),
        "adv256": synth.adversarial_clouds(256, 0),
        "dense1000": synth.make_clouds(1, 1000, 3, "dense"),   # N not a power of two: bs=512, ragged strides
        "dense100": synth.make_clouds(2, 100, 4, "dense", role="template"),
    }
    for name, c in clouds.items():
        xyz = t(c)
        n = c.shape[1]
        m = n // 2
        idx = cops.furthest_point_sampling(xyz, m)
        new_xyz = cops.gather_points(xyz.transpose(1, 2).contiguous(), idx).transpose(1, 2).contiguous()
        out[name + "/xyz"] = c
        out[name + "/fps"] = idx
        for r, ns in ((0.3, 32), (0.7, 16)):
            out[name + "/bq_r%g_ns%d" % (r, ns)] = cops.ball_query(new_xyz, xyz, r, ns)
        out[name + "/knn16"] = cops.knn(new_xyz, min(16, m))
        d2, i3 = cops.three_nn(xyz, new_xyz)
        out[name + "/three_nn_d2"] = d2
        out[name + "/three_nn_idx"] = i3
    save("ops.npz", **out)


def sa_fixture(models):
    from ptt.models.backbones_3d.pointnet2.pointnet2_modules import PointnetSAModuleVotes

    out = {}
    cases = {
        # name: (N, C_in, mlp, npoint, radius, nsample, method, cloud kind)
        "sa1": (1024, 0, [0, 64, 64, 128], 512, 0.3, 32, "fps", "dense"),
        "sa2": (512, 128, [128, 128, 128, 256], 256, 0.5, 32, "sequence", "dense"),
        "sa3_sparse": (256, 256, [256, 128, 128, 256], 128, 0.7, 32, "sequence", "sparse"),
        "box": (128, 257, [257, 256, 256, 256], 64, 0.3, 16, "fps", "dense"),
        "ragged": (200, 5, [5, 24, 40], 50, 0.4, 8, "fps", "sparse"),
    }
    for i, (name, (n, cin, mlp, npoint, radius, ns, method, kind)) in enumerate(cases.items()):
        mod = PointnetSAModuleVotes(mlp=list(mlp), radius=radius, nsample=ns, use_xyz=True, normalize_xyz=True,
                                    sample_method=method).eval()
        synth.load_filled(mod, seed=10 + i)
        xyz = synth.make_clouds(2, n, 20 + i, kind)
        if method == "sequence":  # upstream layers hand over FPS-ordered points; mimic that ordering
            o = cops.furthest_point_sampling(t(xyz), n).long()
            xyz = np.take_along_axis(xyz, o.numpy()[:, :, None], 1)
        feats = synth.features((2, cin, n), seed=30 + i) if cin else None
        new_xyz, new_feats, inds = mod(t(xyz), t(feats) if cin else None, npoint)
        out[name + "/xyz"] = xyz
        if cin:
            out[name + "/features_crc"] = np.uint32(synth.crc(feats))
        out[name + "/new_xyz"] = new_xyz
        out[name + "/new_features"] = new_feats
        out[name + "/inds"] = inds
    save("sa_module.npz", **out)


def transformer_fixture(models):
    from ptt.models.transformer_block import variants

    out = {}
    cases = {
        # name: (class, n, d_points, d_model, k)
        "centroid": ("TransformerBlock", 128, 256, 512, 16),
        "box": ("TransformerBlock", 64, 256, 512, 16),
        "small": ("TransformerBlock", 40, 24, 48, 5),
        "mlp": ("TransformerBlockMLP", 32, 32, 64, 8),
        "offset": ("TransformerBlockOffset", 32, 32, 64, 8),
        "std": ("TransformerBlockSTD", 48, 32, 64, 8),
    }
    for i, (name, (cls, n, dp, dm, k)) in enumerate(cases.items()):
        mod = getattr(variants, cls)(dp, dm, k).eval()
        synth.load_filled(mod, seed=40 + i)
        xyz = synth.make_clouds(2, n, 50 + i, "dense", role="template")
        feats = synth.features((2, n, dp), seed=60 + i)
        res, attn = mod(t(xyz), t(feats))
        out[name + "/xyz"] = xyz
        out[name + "/features_crc"] = np.uint32(synth.crc(feats))
        out[name + "/res"] = res
        out[name + "/attn_head"] = attn[:, :4].contiguous()   # the first 4 tokens only (size)
    # the other registered blocks (transformer_block/__init__.py:7-17), built through the reference's own classes
    from ptt.models.transformer_block import multitransformer
    extra = {
        # name: (class, n, d_points, d_model, k, heads, layers)
        "cosine": ("TransformerBlockCosine", 48, 32, 64, 8, 1, 1),
        "all": ("TransformerBlockALL", 40, 24, 48, 4, 1, 1),
        "cross": ("CrossAttentionBlock", 48, 32, 64, 8, 1, 1),
        "mul": ("MulTransformerBlock", 48, 32, 64, 8, 4, 2),
    }
    for i, (name, (cls, n, dp, dm, k, heads, layers)) in enumerate(extra.items()):
        ctor = multitransformer.MulTransformerBlock if cls == "MulTransformerBlock" else getattr(variants, cls)
        mod = ctor(d_points=dp, d_model=dm, k=k, heads=heads, layers=layers).eval()
        synth.load_filled(mod, seed=80 + i)
        xyz = synth.make_clouds(2, n, 90 + i, "dense", role="template")
        feats = synth.features((2, n, dp), seed=100 + i)
        if cls == "CrossAttentionBlock":
            feats2 = synth.features((2, n, dp), seed=110 + i)
            res, attn = mod(t(xyz), t(feats), t(feats2))
        else:
            res, attn = mod(t(xyz), t(feats))
        out[name + "/xyz"] = xyz
        out[name + "/features_crc"] = np.uint32(synth.crc(feats))
        out[name + "/res"] = res
        out[name + "/attn_head"] = attn[:, :4].contiguous()
    save("transformer.npz", **out)


def hot_path_fixture(models):
    """The bench/smoke frame (oracle.torch_port.hot_path_frame) computed with the REFERENCE modules of
    a full PTT tracker built from tools/cfgs/kitti_models/ptt.yaml."""
    net, cfg = refload.build_tracker(training=False)
    synth.load_filled(net, seed=0)
    out = {}
    for name, kind, ns, nt in (("dense", "dense", 1024, 512), ("sparse", "sparse", 512, 512)):
        search = synth.make_clouds(2, ns, 70, kind)
        template = synth.make_clouds(2, nt, 71, kind, role="template")
        bd = net.backbone_3d({"search_points": t(search), "template_points": t(template)})
        cen = net.centroid_voting_head.transformer_block(
            xyz=bd["search_seeds"], features=bd["search_feats"].transpose(1, 2).contiguous())[0]
        votes_feats = torch.cat([torch.full_like(cen[:, :, :1], 0.5), cen], dim=2).transpose(1, 2).contiguous()
        b_xyz, b_feat, _ = net.box_voting_head.vote_aggregation(
            xyz=bd["search_seeds"], features=votes_feats, npoint=cfg.MODEL.BOX_HEAD.SA_CONFIG.NPOINTS)
        box = net.box_voting_head.transformer_block(xyz=b_xyz, features=b_feat.transpose(1, 2).contiguous())[0]
        out[name + "/search"] = search
        out[name + "/template"] = template
        for k in ("search_seeds", "search_feats", "search_inds", "template_seeds", "template_feats", "template_inds"):
            out[name + "/" + k] = bd[k]
        out[name + "/centroid_feats"] = cen
        out[name + "/box_centers"] = b_xyz
        out[name + "/box_sa_feats"] = b_feat
        out[name + "/box_feats"] = box
    # whole reference model, all modules (for the drop-in `_ext` test and later "next" rows)
    search = synth.make_clouds(2, 1024, 72, "dense")
    template = synth.make_clouds(2, 512, 73, "dense", role="template")
    full = net({"search_points": t(search), "template_points": t(template)})
    out["full/search"] = search
    out["full/template"] = template
    for k in ("search_feats", "template_feats", "cosine_feats", "pred_centroids_cls", "pred_centroids_votes",
              "pred_box_center", "pred_box_data"):
        out["full/" + k] = full[k]
    save("hot_path.npz", **out)


if __name__ == "__main__":
    cops.build()
    models = refload.load()
    ops_fixture()
    sa_fixture(models)
    transformer_fixture(models)
    hot_path_fixture(models)
