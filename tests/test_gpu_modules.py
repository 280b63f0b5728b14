"""GPU parity: the fused SA layer, the transformer block and the whole hot path vs (a) the committed golden
fixtures the REFERENCE's own modules produced (tests/golden/make_golden.py) and (b) the CPU oracle port.

Tolerances: indices bit-exact; fp32 features within 1e-4 (north_star)."""
import numpy as np
import pytest
import torch

from conftest import t
from oracle import cops, torch_port
from ptt_b200 import hotpath, modules, ops, synth
from test_oracle_golden import SA_CASES, TR_CASES, sa_state_dict, transformer_state_dict

pytestmark = pytest.mark.gpu

DEV = "cuda:0"
FP_TOL = dict(rtol=1e-4, atol=1e-4)


def g(a):
    if isinstance(a, np.ndarray):
        a = t(a)
    return a.to(DEV).contiguous()


def filled(sd_template, seed):
    return {k: t(v) for k, v in synth.fill_state_dict(sd_template, seed=seed).items()}


@pytest.mark.parametrize("name", list(SA_CASES))
def test_sa_module_vs_reference_fixture(golden, name):
    gd = golden("sa_module.npz")
    i = list(SA_CASES).index(name)
    n, cin, mlp, npoint, radius, ns, method = SA_CASES[name]
    mod = modules.PointnetSAModuleVotes(mlp=list(mlp), radius=radius, nsample=ns, use_xyz=True, normalize_xyz=True,
                                        sample_method=method)
    mod.load_state_dict(filled(sa_state_dict(mlp), 10 + i))
    mod = mod.to(DEV).eval()
    feats = g(synth.features((2, cin, n), seed=30 + i)) if cin else None
    with torch.no_grad():
        new_xyz, new_feats, inds = mod(g(gd[name + "/xyz"]), feats, npoint)
    assert inds.dtype == torch.int64
    assert np.array_equal(inds.cpu().numpy(), gd[name + "/inds"])
    assert np.array_equal(new_xyz.cpu().numpy(), gd[name + "/new_xyz"])
    np.testing.assert_allclose(new_feats.cpu().numpy(), gd[name + "/new_features"], **FP_TOL)


@pytest.mark.parametrize("name", ["centroid", "box", "small", "offset", "std"])
def test_transformer_vs_reference_fixture(golden, name):
    gd = golden("transformer.npz")
    i = list(TR_CASES).index(name)
    cls, n, dp, dm, k = TR_CASES[name]
    mod = getattr(modules, cls)(dp, dm, k, heads=1, layers=1)
    mod.load_state_dict(filled(transformer_state_dict(cls, dp, dm), 40 + i))
    mod = mod.to(DEV).eval()
    xyz = gd[name + "/xyz"]
    f = synth.features((2, n, dp), seed=60 + i)
    assert synth.crc(f) == int(gd[name + "/features_crc"])
    with torch.no_grad():
        res, attn = mod(g(xyz), g(f))
    np.testing.assert_allclose(res.cpu().numpy(), gd[name + "/res"], **FP_TOL)
    np.testing.assert_allclose(attn[:, :4].cpu().numpy(), gd[name + "/attn_head"], **FP_TOL)


def test_transformer_with_exact_ties_matches_port():
    # duplicate-heavy cloud: kNN ties are resolved lowest-index-first on both sides
    cls, n, dp, dm, k = "TransformerBlock", 64, 32, 64, 16
    sd = filled(transformer_state_dict(cls, dp, dm), 7)
    xyz = synth.make_clouds(3, n, 9, "sparse")
    f = synth.features((3, n, dp), seed=9)
    want, want_attn = torch_port.transformer_block(sd, t(xyz), t(f), k)
    packed = ops.PackedTransformer({kk: g(v) for kk, v in sd.items()}, k)
    got, attn = ops.transformer_block_fwd(packed, g(xyz), g(f), want_attn=True)
    np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), **FP_TOL)
    np.testing.assert_allclose(attn.cpu().numpy(), want_attn.numpy(), **FP_TOL)


def test_sa_layer_random_shapes_vs_port():
    cases = [(300, 7, [7, 16, 32], 100, 0.5, 12, "fps"), (64, 0, [0, 8], 64, 0.9, 3, "fps"),
             (512, 64, [64, 64, 64, 64, 96], 128, 0.4, 32, "sequence"),
             # shapes the fused tcgen05 kernel takes: ns in {8, 16}, partial last tile, with and without features
             (200, 10, [10, 64, 64, 128], 50, 0.5, 16, "fps"), (100, 0, [0, 128, 128, 256], 33, 0.6, 8, "fps"),
             (256, 128, [128, 128, 128, 256], 77, 0.5, 32, "fps"), (300, 0, [0, 64, 64, 128], 300, 0.3, 4, "fps")]
    for i, (n, cin, mlp, npoint, radius, ns, method) in enumerate(cases):
        sd = filled(sa_state_dict(mlp), 90 + i)
        xyz = synth.make_clouds(2, n, 91 + i, "sparse" if i == 0 else "dense")
        feats = synth.features((2, cin, n), seed=92 + i) if cin else None
        w_xyz, w_feats, w_inds = torch_port.sa_module_votes(sd, t(xyz), t(feats) if cin else None, npoint, radius, ns,
                                                             method, True, True)
        mod = modules.PointnetSAModuleVotes(mlp=list(mlp), radius=radius, nsample=ns, normalize_xyz=True,
                                            sample_method=method)
        mod.load_state_dict(sd)
        mod = mod.to(DEV).eval()
        with torch.no_grad():
            new_xyz, new_feats, inds = mod(g(xyz), g(feats) if cin else None, npoint)
        assert np.array_equal(inds.cpu().numpy(), w_inds.numpy())
        assert np.array_equal(new_xyz.cpu().numpy(), w_xyz.numpy())
        np.testing.assert_allclose(new_feats.cpu().numpy(), w_feats.numpy(), **FP_TOL)


def test_sa_module_training_mode_matches_port_forward_and_grads():
    # training mode runs torch's own Conv2d/BatchNorm (as the reference does); keep cuDNN in true fp32 for the comparison
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    n, cin, mlp, npoint, radius, ns = 128, 16, [16, 32, 48], 40, 0.6, 8
    sd = filled(sa_state_dict(mlp), 123)
    xyz = synth.make_clouds(3, n, 124, "dense", role="template")
    feats = synth.features((3, cin, n), seed=125)
    # oracle side: CPU autograd through the port (batch statistics: training=True)
    sd_cpu = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd.items()}
    xyz_c = t(xyz).clone().requires_grad_(True)
    f_c = t(feats).clone().requires_grad_(True)
    w_xyz, w_feats, _ = _port_sa_training(sd_cpu, xyz_c, f_c, npoint, radius, ns)
    (w_feats.square().sum() + w_xyz.sum()).backward()

    mod = modules.PointnetSAModuleVotes(mlp=list(mlp), radius=radius, nsample=ns, normalize_xyz=True, sample_method="fps")
    mod.load_state_dict(sd)
    mod = mod.to(DEV).train()
    xyz_g = g(xyz).requires_grad_(True)
    f_g = g(feats).requires_grad_(True)
    new_xyz, new_feats, _ = mod(xyz_g, f_g, npoint)
    (new_feats.square().sum() + new_xyz.sum()).backward()
    np.testing.assert_allclose(new_feats.detach().cpu().numpy(), w_feats.detach().numpy(), rtol=1e-3, atol=1e-3)
    np.testing.assert_allclose(f_g.grad.cpu().numpy(), f_c.grad.numpy(), rtol=1e-3, atol=1e-3)
    np.testing.assert_allclose(xyz_g.grad.cpu().numpy(), xyz_c.grad.numpy(), rtol=1e-3, atol=2e-3)
    gw = mod.mlp_module.layer0.conv.weight.grad.cpu().numpy()
    np.testing.assert_allclose(gw, sd_cpu["mlp_module.layer0.conv.weight"].grad.numpy(), rtol=1e-3, atol=2e-3)


def _port_sa_training(sd, xyz, feats, npoint, radius, ns):
    """The port's SA layer with differentiable gathers (torch indexing) for the CPU side of the grad test."""
    import torch.nn.functional as F
    inds = cops.furthest_point_sampling(xyz.detach().contiguous(), npoint).long()
    new_xyz = torch.gather(xyz, 1, inds[:, :, None].expand(-1, -1, 3))
    idx = cops.ball_query(new_xyz.detach().contiguous(), xyz.detach().contiguous(), radius, ns).long()
    B, M, K = idx.shape
    flat = idx.reshape(B, 1, M * K)
    gx = torch.gather(xyz.transpose(1, 2), 2, flat.expand(-1, 3, -1)).reshape(B, 3, M, K)
    gx = (gx - new_xyz.transpose(1, 2).unsqueeze(-1)) / radius
    gf = torch.gather(feats, 2, flat.expand(-1, feats.shape[1], -1)).reshape(B, -1, M, K)
    y = torch.cat([gx, gf], 1)
    i = 0
    while "mlp_module.layer%d.conv.weight" % i in sd:
        p = "mlp_module.layer%d." % i
        y = F.conv2d(y, sd[p + "conv.weight"])
        y = F.batch_norm(y, None, None, sd[p + "normlayer.bn.weight"], sd[p + "normlayer.bn.bias"], training=True)
        y = F.relu(y)
        i += 1
    return new_xyz, y.max(dim=3)[0], inds


@pytest.mark.parametrize("name", ["dense", "sparse"])
def test_hot_path_vs_reference_fixture(golden, name):
    gd = golden("hot_path.npz")
    sd = synth.hot_path_state_dict(0)
    hp = hotpath.HotPath(sd, device=DEV)
    out = hp(g(gd[name + "/search"]), g(gd[name + "/template"]))
    torch.cuda.synchronize()
    for k in ("search_inds", "template_inds"):
        assert np.array_equal(out[k].cpu().numpy(), gd[name + "/" + k]), k
    for k in ("search_seeds", "template_seeds", "box_centers"):
        assert np.array_equal(out[k].cpu().numpy(), gd[name + "/" + k]), k
    for k in ("search_feats", "template_feats", "centroid_feats", "box_sa_feats", "box_feats"):
        np.testing.assert_allclose(out[k].cpu().numpy(), gd[name + "/" + k], err_msg=k, **FP_TOL)


def test_hot_path_full_batch_properties():
    """BASELINE config 2 size (B=48, N=1024/512): size-independent properties + a per-frame spot check
    against the oracle on 2 of the 48 frames."""
    B = 48
    sd = synth.hot_path_state_dict(0)
    hp = hotpath.HotPath(sd, device=DEV)
    search = synth.make_clouds(B, 1024, 500, "dense")
    template = synth.make_clouds(B, 512, 501, "dense", role="template")
    out = hp(g(search), g(template))
    torch.cuda.synchronize()
    inds = out["search_inds"].cpu().numpy()
    assert inds.shape == (B, 128) and (inds[:, 0] == 0).all()
    assert all(len(set(r.tolist())) == 128 for r in inds)          # distinct points -> distinct samples
    seeds = out["search_seeds"].cpu().numpy()
    assert np.array_equal(seeds, np.take_along_axis(search, inds[:, :, None], 1))
    for k, v in out.items():
        assert torch.isfinite(v.float()).all(), k
    # batch independence: frames 5 and 41 alone give the same answer as inside the batch of 48
    sub = hp(g(search[[5, 41]]), g(template[[5, 41]]))
    for k in out:
        a, b = out[k][[5, 41]].cpu().numpy(), sub[k].cpu().numpy()
        if a.dtype.kind == "i":
            assert np.array_equal(a, b), k
        else:
            np.testing.assert_allclose(a, b, rtol=1e-5, atol=1e-5, err_msg=k)
    want = torch_port.hot_path_frame(sd, t(search[[5, 41]]), t(template[[5, 41]]))
    for k in ("search_inds", "template_inds"):
        assert np.array_equal(sub[k].cpu().numpy(), want[k].numpy()), k
    for k in ("search_feats", "template_feats", "centroid_feats", "box_sa_feats", "box_feats"):
        np.testing.assert_allclose(sub[k].cpu().numpy(), want[k].numpy(), err_msg=k, **FP_TOL)


@pytest.mark.parametrize("cluster", [1, 2, 4, -2])
@pytest.mark.parametrize("shape", [(3, 128, 256, 512, 16), (5, 64, 64, 128, 8), (2, 100, 32, 256, 4), (1, 32, 128, 64, 32),
                                   (1, 33, 32, 128, 4), (1, 35, 64, 512, 2), (3, 21, 16, 256, 1)])
def test_transformer_cluster_sizes_agree_with_port(cluster, shape):
    """Every thread-block-cluster size of the tcgen05 transformer passes (weights multicast to 1, 2 or 4 CTAs),
    including tile counts that are not a multiple of the cluster size (dummy tiles) and pair counts that are not a
    multiple of 32 (the guarded last block of the transposed epilogues), against the CPU port."""
    import ctypes
    from ptt_b200 import _lib
    B, n, dp, dm, k = shape
    sd = filled(transformer_state_dict("TransformerBlock", dp, dm), 300 + n)
    xyz = synth.make_clouds(B, n, 301 + n, "dense", role="template")
    f = synth.features((B, n, dp), seed=302 + n)
    want, want_attn = torch_port.transformer_block(sd, t(xyz), t(f), k)
    packed = ops.PackedTransformer({kk: g(v) for kk, v in sd.items()}, k)
    setter = _lib.lib().ptt_debug_set_cluster
    setter.argtypes = [ctypes.c_int]
    setter(cluster)
    try:
        got, attn = ops.transformer_block_fwd(packed, g(xyz), g(f), want_attn=True)
        torch.cuda.synchronize()
    finally:
        setter(0)
    np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), **FP_TOL)
    np.testing.assert_allclose(attn.cpu().numpy(), want_attn.numpy(), **FP_TOL)


def test_hot_path_graph_replay_matches_eager_and_host_api():
    sd = synth.hot_path_state_dict(0)
    hp = hotpath.HotPath(sd, device=DEV)
    outs = []
    for seed in (600, 601, 602):
        search = g(synth.make_clouds(4, 1024, seed, "dense"))
        template = g(synth.make_clouds(4, 512, seed + 50, "dense", role="template"))
        eager = {k: v.clone() for k, v in hp(search, template).items()}
        replay = {k: v.clone() for k, v in hp.forward_graph(search, template).items()}
        torch.cuda.synchronize()
        for k in eager:
            assert torch.equal(eager[k], replay[k]), k          # same kernels, same order: bit-identical
        host = hp.forward_host(search.cpu().pin_memory(), template.cpu().pin_memory(), outputs="all")
        assert set(host) == set(hp.HOST_KEYS)
        for k in host:
            assert torch.equal(eager[k].cpu(), host[k]), k
        dflt = hp.forward_host(search.cpu().pin_memory(), template.cpu().pin_memory())     # default selection
        assert list(dflt) == ["box_feats"] and torch.equal(dflt["box_feats"], eager["box_feats"].cpu())
        outs.append(eager["box_feats"])
    assert not torch.equal(outs[0], outs[1])
    with pytest.raises(KeyError):
        hp.forward_host(search.cpu().pin_memory(), template.cpu().pin_memory(), outputs=("no_such_output",))


def test_graphs_of_different_shapes_keep_their_own_workspaces():
    """A graph captured for a small batch must replay correctly after a LARGER shape has been run (eagerly and as a
    graph) on the same HotPath: every captured graph owns its workspaces (a re-allocated shared workspace
    left the first graph replaying into freed memory)."""
    sd = synth.hot_path_state_dict(0)
    hp = hotpath.HotPath(sd, device=DEV)
    small = (g(synth.make_clouds(2, 1024, 610, "dense")), g(synth.make_clouds(2, 512, 611, "dense", role="template")))
    big = (g(synth.make_clouds(12, 1024, 612, "dense")), g(synth.make_clouds(12, 512, 613, "dense", role="template")))
    first = {k: v.clone() for k, v in hp.forward_graph(*small).items()}
    hp(*big)                                                   # eager, larger: grows the eager-scope workspaces
    big_out = {k: v.clone() for k, v in hp.forward_graph(*big).items()}
    junk = [torch.full((1 << 22,), float("nan"), device=DEV) for _ in range(8)]     # recycle whatever was freed
    again = hp.forward_graph(*small)
    torch.cuda.synchronize()
    for k in first:
        assert torch.equal(first[k], again[k]), k
    eager_big = hp(*big)
    for k in big_out:
        assert torch.equal(big_out[k], eager_big[k]), k
    del junk


def test_host_pipeline_matches_synchronous_api():
    """Results come back in submission order and -- WITHOUT cloning them -- survive the step that the returning push
    itself enqueued on the same slot (push() used to hand out buffers the in-flight step was overwriting)."""
    sd = synth.hot_path_state_dict(0)
    hp = hotpath.HotPath(sd, device=DEV)
    for outputs in ("all", None, ("box_feats", "search_inds")):
        pipe = hotpath.HostPipeline(sd, device=DEV, depth=2, outputs=outputs)
        batches = [(t(synth.make_clouds(3, 1024, 700 + i, "dense")).pin_memory(),
                    t(synth.make_clouds(3, 512, 750 + i, "dense", role="template")).pin_memory()) for i in range(6)]
        want = [{k: v.clone() for k, v in hp.forward_host(s, tm, outputs=outputs).items()} for s, tm in batches]
        n_got = 0
        for s, tm in batches:
            r = pipe.push(s, tm)
            if r is not None:
                torch.cuda.synchronize()        # the step just submitted on r's slot has finished: r must be intact
                assert set(r) == set(want[n_got])
                for k in r:
                    assert torch.equal(r[k], want[n_got][k]), (outputs, n_got, k)
                n_got += 1
        for r in pipe.drain():
            for k in r:
                assert torch.equal(r[k], want[n_got][k]), (outputs, n_got, k)
            n_got += 1
        assert n_got == len(want)
        for to_host in (False,):                # device-resident results follow the same rotation
            r0 = None
            for i, (s, tm) in enumerate(batches[:4]):
                r = pipe.push(s.to(DEV), tm.to(DEV), to_host=to_host)
                if r is not None and r0 is None:
                    torch.cuda.synchronize()
                    r0 = {k: v.clone() for k, v in r.items()}
                    for k in r0:
                        assert r[k].is_cuda and torch.equal(r[k].cpu(), want[0][k]), k
            pipe.drain()


def test_eval_mode_with_grad_enabled_builds_the_graph():
    """eval() + grad enabled (frozen-BatchNorm fine-tuning, saliency) must not take the fused path, whose
    outputs carry no autograd graph; under no_grad the fused path is kept and both agree."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    sa = modules.PointnetSAModuleVotes(mlp=[0, 64, 64, 128], radius=0.3, nsample=32, normalize_xyz=True).to(DEV).eval()
    tr = modules.TransformerBlock(128, 64, 8).to(DEV).eval()
    xyz = g(synth.make_clouds(2, 256, 31, "dense"))
    with torch.no_grad():
        nx0, f0, _ = sa(xyz, None, 64)
        r0, _ = tr(nx0, f0.transpose(1, 2).contiguous())
    assert not f0.requires_grad
    nx, f, _ = sa(xyz, None, 64)                        # grad mode on, parameters require grad
    r, _ = tr(nx, f.transpose(1, 2).contiguous())
    assert f.requires_grad and r.requires_grad
    r.square().mean().backward()
    assert sa.mlp_module.layer0.conv.weight.grad is not None and tr.fc_gamma[0].weight.grad is not None
    np.testing.assert_allclose(f.detach().cpu().numpy(), f0.cpu().numpy(), **FP_TOL)
    np.testing.assert_allclose(r.detach().cpu().numpy(), r0.cpu().numpy(), **FP_TOL)
    # frozen parameters + grad mode on + inputs without grad -> nobody can ask for a gradient: fused path again
    for p in list(sa.parameters()) + list(tr.parameters()):
        p.requires_grad_(False)
    _, f2, _ = sa(xyz, None, 64)
    assert not f2.requires_grad and torch.equal(f2, f0)


def test_packed_weights_follow_in_place_updates():
    """an in-place parameter update in eval mode (optimizer step with frozen BN, EMA swap through copy_)
    must not leave a stale packed image in use."""
    tr = modules.TransformerBlock(32, 64, 4).to(DEV).eval()
    sa = modules.PointnetSAModuleVotes(mlp=[0, 16, 32], radius=0.5, nsample=8, normalize_xyz=True).to(DEV).eval()
    xyz = g(synth.make_clouds(2, 64, 33, "dense", role="template"))
    feats = g(synth.features((2, 64, 32), seed=34))
    with torch.no_grad():
        a0, _ = tr(xyz, feats)
        _, s0, _ = sa(xyz, None, 16)
        tr.fc2.weight.mul_(2.0)
        sa.mlp_module.layer1.normlayer.bn.running_var.fill_(4.0)
        a1, _ = tr(xyz, feats)
        _, s1, _ = sa(xyz, None, 16)
        fresh = modules.TransformerBlock(32, 64, 4).to(DEV).eval()
        fresh.load_state_dict(tr.state_dict())
        assert torch.equal(a1, fresh(xyz, feats)[0]) and not torch.equal(a0, a1)
        fresh_sa = modules.PointnetSAModuleVotes(mlp=[0, 16, 32], radius=0.5, nsample=8, normalize_xyz=True).to(DEV).eval()
        fresh_sa.load_state_dict(sa.state_dict())
        assert torch.equal(s1, fresh_sa(xyz, None, 16)[1]) and not torch.equal(s0, s1)
        tr.fc2.bias.data.add_(1.0)              # .data bypasses the version counter: explicit invalidate()
        tr.invalidate()
        a2, _ = tr(xyz, feats)
        np.testing.assert_allclose(a2.cpu().numpy(), (a1 + 1.0).cpu().numpy(), **FP_TOL)


@pytest.mark.parametrize("shape", [(3, 128, 256, 512), (2, 100, 32, 64), (1, 1024, 64, 128), (4, 37, 24, 48)])
def test_transformer_std_vs_port(shape):
    """Dense n x n attention block (the batched tcgen05 contractions) against the CPU port, incl. N = 1024 tokens,
    token counts that are not multiples of 64 / 4 and a d_model that is not a multiple of 64."""
    B, n, dp, dm = shape
    sd = filled(transformer_state_dict("TransformerBlockSTD", dp, dm), 400 + n)
    xyz = synth.make_clouds(B, n, 401 + n, "dense", role="template")
    f = synth.features((B, n, dp), seed=402 + n)
    want, want_attn = torch_port.transformer_block_std(sd, t(xyz), t(f))
    packed = ops.PackedTransformerSTD({kk: g(v) for kk, v in sd.items()})
    got, attn = ops.transformer_std_fwd(packed, g(xyz), g(f))
    np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), **FP_TOL)
    np.testing.assert_allclose(attn.cpu().numpy(), want_attn.numpy(), **FP_TOL)


# ------------------------------------------------------------------------------------------------------------------
# the similarity module and the heads' Conv1d stacks -> the whole tracker forward
# ------------------------------------------------------------------------------------------------------------------
def _sub(sd, prefix):
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}


@pytest.mark.parametrize("shape", [(3, 64, 128, 256), (2, 16, 40, 256), (2, 24, 33, 256), (1, 128, 8, 256)])
def test_cosine_fusion_vs_port(shape):
    """CosineSimAug: fused path (n1 a power of two) and the generic path (n1 = 24), against the port."""
    B, n1, n2, f = shape
    sd = {k: v for k, v in synth.full_model_state_dict(3).items() if k.startswith("similarity_module.")}
    sd = _sub(sd, "similarity_module.")
    s_feat = synth.features((B, n2, f), seed=900 + n1)
    t_feat = synth.features((B, n1, f), seed=901 + n1)
    t_xyz = synth.make_clouds(B, n1, 902, "dense", role="template")
    cf = ops.PackedCosineFusion({k: g(v) for k, v in sd.items()})
    got = cf(g(s_feat), g(t_feat), g(t_xyz))
    want = torch_port.cosine_sim_aug(sd, t(s_feat).transpose(1, 2).contiguous(), t(t_feat).transpose(1, 2).contiguous(), t(t_xyz))
    np.testing.assert_allclose(got.cpu().numpy(), want.transpose(1, 2).numpy(), **FP_TOL)


def _check_full(out, want, sd, cfg=None):
    """Stage-by-stage comparison of HotPath.forward_full with the port.  The box head starts with FPS over the
    PREDICTED votes, where a 1e-6 difference may legitimately flip a pick, so its reference is the port's box head
    run on the votes the GPU produced (identical FPS input); when the picks agree with the port's own end-to-end run
    as well (they do on the committed cases) the final boxes are compared directly too."""
    for k in ("search_inds", "template_inds"):
        assert np.array_equal(out[k].cpu().numpy(), want[k].numpy()), k
    for k in ("search_feats", "template_feats", "cosine_feats", "pred_centroids_cls", "pred_centroids_votes", "votes_feats"):
        np.testing.assert_allclose(out[k].cpu().numpy(), want[k].numpy(), err_msg=k, **FP_TOL)
    np.testing.assert_allclose(out["centroid_feats"].cpu().numpy(), want["centroid_feats"].numpy(), **FP_TOL)
    c = dict(knn=16, box_npoint=64, box_radius=0.3, box_nsample=16)
    c.update(cfg or {})
    b_xyz, box_data, b_feat, box = torch_port.box_voting_head(_sub(sd, "box_voting_head."), out["pred_centroids_votes"].cpu(),
                                                              out["votes_feats"].cpu(), c)
    assert np.array_equal(out["pred_box_center"].cpu().numpy(), b_xyz.numpy())
    np.testing.assert_allclose(out["box_sa_feats"].cpu().numpy(), b_feat.transpose(1, 2).numpy(), **FP_TOL)
    np.testing.assert_allclose(out["box_feats"].cpu().numpy(), box.numpy(), **FP_TOL)
    np.testing.assert_allclose(out["pred_box_data"].cpu().numpy(), box_data.numpy(), **FP_TOL)
    same_picks = np.allclose(out["pred_box_center"].cpu().numpy(), want["pred_box_center"].numpy(), rtol=1e-5, atol=1e-4)
    if same_picks:
        np.testing.assert_allclose(out["pred_box_data"].cpu().numpy(), want["pred_box_data"].numpy(), **FP_TOL)
    return same_picks


def test_full_model_vs_reference_fixture(golden):
    """The whole tracker forward on the GPU against what the REFERENCE's own PTT model produced (full/*)."""
    gd = golden("hot_path.npz")
    sd = synth.full_model_state_dict(0)
    hp = hotpath.HotPath(sd, device=DEV)
    out = hp.forward_full(g(gd["full/search"]), g(gd["full/template"]))
    torch.cuda.synchronize()
    for k in ("search_feats", "template_feats", "cosine_feats", "pred_centroids_cls", "pred_centroids_votes"):
        np.testing.assert_allclose(out[k].cpu().numpy(), gd["full/" + k], err_msg=k, **FP_TOL)
    want = torch_port.full_model_frame(sd, t(gd["full/search"]), t(gd["full/template"]))
    same = _check_full(out, want, sd)
    assert same, "box-centre FPS picks differ from the reference run"
    # the proposal selection of the evaluation loop (eval_tracking_utils.py:268-270) done on the device
    est = out["pred_box_data"].cpu().numpy()
    pick = est[:, :, 4].argmax(1)
    assert np.array_equal(out["best_idx"].cpu().numpy(), pick)
    assert np.array_equal(out["best_box"].cpu().numpy(), est[np.arange(est.shape[0]), pick])
    np.testing.assert_allclose(out["pred_box_center"].cpu().numpy(), gd["full/pred_box_center"], rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(out["pred_box_data"].cpu().numpy(), gd["full/pred_box_data"], **FP_TOL)


def test_full_model_sparse_and_small_vs_port():
    sd = synth.full_model_state_dict(1)
    hp = hotpath.HotPath(sd, device=DEV)
    search, template = synth.make_clouds(3, 512, 910, "sparse"), synth.make_clouds(3, 512, 911, "sparse", role="template")
    out = hp.forward_full(g(search), g(template))
    torch.cuda.synchronize()
    _check_full(out, torch_port.full_model_frame(sd, t(search), t(template)), sd)
    cfg = dict(npoints_search=(128, 64, 32), npoints_template=(64, 32, 16), box_npoint=16)
    hp = hotpath.HotPath(sd, cfg=cfg, device=DEV)
    search, template = synth.make_clouds(2, 256, 912, "dense"), synth.make_clouds(2, 128, 913, "dense", role="template")
    out = hp.forward_full(g(search), g(template))
    torch.cuda.synchronize()
    _check_full(out, torch_port.full_model_frame(sd, t(search), t(template), cfg), sd, cfg)


# ------------------------------------------------------------------------------------------------------------------
# a10 / N4: the other blocks of transformer_block.__all__
# ------------------------------------------------------------------------------------------------------------------
from test_oracle_golden import EXTRA_TR_CASES, extra_transformer_state_dict, run_extra_port  # noqa: E402


def _build_extra(cls, dp, dm, k, heads, layers, sd):
    mod = modules.REGISTRY[cls](d_points=dp, d_model=dm, k=k, heads=heads, layers=layers)
    mod.load_state_dict(sd)
    return mod.to(DEV).eval()


@pytest.mark.parametrize("name", list(EXTRA_TR_CASES))
def test_secondary_blocks_vs_reference_fixture(golden, name):
    gd = golden("transformer.npz")
    i = list(EXTRA_TR_CASES).index(name)
    cls, n, dp, dm, k, heads, layers = EXTRA_TR_CASES[name]
    sd = filled(extra_transformer_state_dict(cls, dp, dm, heads, layers), 80 + i)
    mod = _build_extra(cls, dp, dm, k, heads, layers, sd)
    f = synth.features((2, n, dp), seed=100 + i)
    args = [g(gd[name + "/xyz"]), g(f)]
    if cls == "CrossAttentionBlock":
        args.append(g(synth.features((2, n, dp), seed=110 + i)))
    with torch.no_grad():
        res, attn = mod(*args)
    np.testing.assert_allclose(res.cpu().numpy(), gd[name + "/res"], **FP_TOL)
    np.testing.assert_allclose(attn[:, :4].cpu().numpy(), gd[name + "/attn_head"], **FP_TOL)


@pytest.mark.parametrize("case", [("TransformerBlockCosine", 3, 128, 256, 512, 16, 1, 1), ("TransformerBlockCosine", 2, 50, 64, 128, 4, 1, 1),
                                  ("CrossAttentionBlock", 3, 64, 256, 512, 16, 1, 1), ("MulTransformerBlock", 2, 128, 256, 512, 16, 8, 1),
                                  ("MulTransformerBlock", 2, 40, 32, 256, 8, 1, 2), ("TransformerBlockALL", 3, 128, 256, 512, 16, 1, 1),
                                  ("TransformerBlockALL", 2, 33, 20, 40, 3, 1, 1), ("CrossAttentionBlock", 2, 30, 24, 40, 5, 1, 1)])
def test_secondary_blocks_vs_port(case):
    """The production sizes (d_model 512: CTA-pair tcgen05 passes) and ragged / generic-path sizes against the port."""
    cls, B, n, dp, dm, k, heads, layers = case
    sd = filled(extra_transformer_state_dict(cls, dp, dm, heads, layers), 200 + n)
    mod = _build_extra(cls, dp, dm, k, heads, layers, sd)
    xyz = synth.make_clouds(B, n, 300 + n, "dense", role="template")
    f = synth.features((B, n, dp), seed=310 + n)
    f2 = synth.features((B, n, dp), seed=320 + n) if cls == "CrossAttentionBlock" else None
    with torch.no_grad():
        res, attn = mod(*([g(xyz), g(f)] + ([g(f2)] if f2 is not None else [])))
    want_res, want_attn = run_extra_port(cls, sd, t(xyz), t(f), k, heads, t(f2) if f2 is not None else None)
    np.testing.assert_allclose(res.cpu().numpy(), want_res.numpy(), **FP_TOL)
    np.testing.assert_allclose(attn.cpu().numpy(), want_attn.numpy(), **FP_TOL)


@pytest.mark.parametrize("shape", [(2, 64, 256, 512, 16), (2, 32, 32, 64, 8)])
def test_mlp_and_backbone_blocks_vs_port(shape):
    B, n, dp, dm, k = shape
    sd = filled(transformer_state_dict("TransformerBlockMLP", dp, dm), 400 + n)
    mod = modules.TransformerBlockMLP(dp, dm, k)
    mod.load_state_dict(sd)
    mod = mod.to(DEV).eval()
    xyz = synth.make_clouds(B, n, 410 + n, "dense", role="template")
    f = synth.features((B, n, dp), seed=420 + n)
    with torch.no_grad():
        res, attn = mod(g(xyz), g(f))
    want_res, want_attn = torch_port.transformer_block(sd, t(xyz), t(f), k, variant="TransformerBlockMLP")
    np.testing.assert_allclose(res.cpu().numpy(), want_res.numpy(), **FP_TOL)
    np.testing.assert_allclose(attn.cpu().numpy(), want_attn.numpy(), **FP_TOL)
    # Backbone: neighbourhoods supplied by the caller (here: a ball query over the points themselves)
    sd = filled(transformer_state_dict("TransformerBlock", dp, dm), 430 + n)
    bb = modules.TransformerBlockBackbone(dp, dm, k)
    bb.load_state_dict(sd)
    bb = bb.to(DEV).eval()
    idx = cops.ball_query(t(xyz), t(xyz), 0.6, k)
    grouped = cops.group_points(t(xyz).transpose(1, 2).contiguous(), idx)
    with torch.no_grad():
        got = bb(g(xyz), g(grouped), g(idx), g(f))
    want = torch_port.transformer_block_backbone(sd, t(xyz), grouped, idx, t(f))
    np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), **FP_TOL)


def test_hot_path_net_state_dict_drives_hotpath_and_trains():
    """ptt_b200.train.HotPathNet: (a) its state_dict has exactly the hot-path keys, so in eval mode it agrees with
    HotPath built from the same parameters; (b) one SGD step in train mode gives finite gradients for every parameter."""
    from ptt_b200 import train
    net = train.HotPathNet()
    sd = synth.hot_path_state_dict(0)
    assert set(net.state_dict().keys()) == set(sd.keys())
    net.load_state_dict(sd)
    net = net.to(DEV)
    search, template = g(synth.make_clouds(2, 1024, 950, "dense")), g(synth.make_clouds(2, 512, 951, "dense", role="template"))
    net.eval()
    # cov_final is a torch Conv1d here: torch lets cuDNN use TF32 for convolutions by default (1e-3 noise); fp32 for the check
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            got = net(search, template)
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    want = hotpath.HotPath(sd, device=DEV)(search, template)
    torch.cuda.synchronize()
    np.testing.assert_allclose(got["box_feats"].cpu().numpy(), want["box_feats"].cpu().numpy(), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(got["search_feats"].cpu().numpy(), want["search_feats"].cpu().numpy(), rtol=1e-4, atol=1e-4)
    net.train()
    out = net(search, template)
    loss = sum((v ** 2).mean() for v in out.values())
    loss.backward()
    for name, p in net.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), name
