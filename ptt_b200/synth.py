"""Synthetic inputs and deterministic parameter fills shared by tests, bench.py and the oracle.

The recipes are frozen: committed golden fixtures depend on them bit for bit, so a
change here invalidates tests/golden/*.npz (regenerate with tests/golden/make_golden.py).
Everything is drawn from numpy's legacy RandomState, which is stable across numpy versions.

Geometry follows the reference's data layer: clouds are expressed in the box-centred object frame
(ptt/datasets/kitti/kitti_tracking_utils.py:312-320), so they straddle the origin; sparse clouds
are resampled WITH replacement to the fixed size (regularize_pc, kitti_tracking_utils.py:342-367),
which yields exact duplicate points, and degenerate crops become all-zero clouds (:359-360).
"""
import zlib

import numpy as np

CAR = (3.9, 1.6, 1.56)          # l, w, h  [m]
PEDESTRIAN = (0.8, 0.6, 1.73)


def _search_half_extents(dims):
    l, w, h = dims
    m = 0.6 * l
    return np.array([1.25 * l / 2 + m, 1.25 * w / 2 + m, 1.25 * h / 2 + m], dtype=np.float64)


def _object_points(rs, n, dims, offset):
    """Uniform on the top and the four side faces of the box (the bottom is never seen)."""
    l, w, h = dims
    areas = np.array([l * w, l * h, l * h, w * h, w * h])
    face = rs.choice(5, size=n, p=areas / areas.sum())
    u = rs.uniform(-0.5, 0.5, size=n)
    v = rs.uniform(-0.5, 0.5, size=n)
    pts = np.empty((n, 3))
    top = face == 0
    pts[top] = np.stack([u[top] * l, v[top] * w, np.full(top.sum(), h / 2)], 1)
    for f, sgn in ((1, 1.0), (2, -1.0)):
        m = face == f
        pts[m] = np.stack([u[m] * l, np.full(m.sum(), sgn * w / 2), v[m] * h], 1)
    for f, sgn in ((3, 1.0), (4, -1.0)):
        m = face == f
        pts[m] = np.stack([np.full(m.sum(), sgn * l / 2), u[m] * w, v[m] * h], 1)
    return pts + offset


def _scene_points(rs, n, dims, half, object_fraction):
    offset = rs.normal(0.0, [0.3, 0.3, 0.05])
    n_obj = int(round(n * object_fraction))
    obj = _object_points(rs, n_obj, dims, offset)
    gx = rs.uniform(-half[0], half[0], size=n - n_obj)
    gy = rs.uniform(-half[1], half[1], size=n - n_obj)
    ground = np.stack([gx, gy, np.full(n - n_obj, -dims[2] / 2)], 1)
    pts = np.concatenate([obj, ground], 0)
    pts = pts[rs.permutation(n)]
    pts += rs.normal(0.0, 0.02, size=pts.shape)
    return np.clip(pts, -half, half)


def make_clouds(batch, n, seed, kind="dense", role="search", dims=CAR):
    """(batch, n, 3) float32 clouds.  kind: 'dense' | 'sparse';  role: 'search' | 'template'."""
    rs = np.random.RandomState((seed * 1000003 + n * 31 + (0 if role == "search" else 7)) % (2 ** 31))
    if role == "search":
        half, frac = _search_half_extents(dims), 0.5
    else:
        half, frac = 1.25 * np.asarray(dims) / 2, 0.9
    out = np.zeros((batch, n, 3), dtype=np.float32)
    for b in range(batch):
        if kind == "dense":
            out[b] = _scene_points(rs, n, dims, half, frac).astype(np.float32)
        elif kind == "sparse":
            if rs.uniform() < 0.01:
                continue  # all-zero cloud
            p = int(rs.randint(21, 201))
            base = _scene_points(rs, p, dims, half, frac).astype(np.float32)
            out[b] = base[rs.randint(0, p, size=n)]
        else:
            raise ValueError(kind)
    return out


def adversarial_clouds(n, seed=0):
    """Edge cases of SURVEY.md 8(c): a (6, n, 3) float32 stack of
    [30 distinct points resampled, all zero, points inside the FPS origin ball, points exactly on a
    0.3 radius shell, a regular grid (many exact distance ties), all points identical]."""
    rs = np.random.RandomState(seed + 77)
    out = np.zeros((6, n, 3), dtype=np.float32)
    base = _scene_points(rs, 30, CAR, _search_half_extents(CAR), 0.5).astype(np.float32)
    out[0] = base[rs.randint(0, 30, size=n)]
    # out[1] stays zero
    pts = _scene_points(rs, n, CAR, _search_half_extents(CAR), 0.5).astype(np.float32)
    near = rs.uniform(-0.03, 0.03, size=(n // 4, 3)).astype(np.float32)   # |p|^2 around the 1e-3 threshold
    pts[rs.permutation(n)[: n // 4]] = near
    out[2] = pts
    d = rs.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    shell = (0.3 * d).astype(np.float32)
    shell[0] = 0.0
    out[3] = shell
    g = int(np.ceil(n ** (1 / 3)))
    grid = np.stack(np.meshgrid(*[np.arange(g)] * 3, indexing="ij"), -1).reshape(-1, 3)[:n]
    out[4] = (grid * 0.25 - 0.25 * g / 2 + 0.125).astype(np.float32)
    out[5] = np.float32([0.7, -0.2, 0.1])
    return out


def _key_seed(seed, key):
    return (zlib.crc32(key.encode()) ^ (seed * 2654435761)) % (2 ** 32)


def fill_state_dict(state_dict, seed=0):
    """Deterministic, key-addressed values for every tensor of a state_dict (returns a new dict of
    numpy arrays).  Independent of construction order and of torch's RNG, so the reference module
    (in this container) and the B200 module (on the GPU box) receive identical parameters."""
    out = {}
    for key, t in state_dict.items():
        shape = tuple(t) if isinstance(t, (tuple, list)) else tuple(t.shape)
        rs = np.random.RandomState(_key_seed(seed, key))
        leaf = key.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked":
            out[key] = np.zeros(shape, dtype=np.int64)
        elif leaf == "running_var":
            out[key] = rs.uniform(0.5, 1.5, size=shape).astype(np.float32)
        elif leaf == "running_mean":
            out[key] = (0.1 * rs.standard_normal(shape)).astype(np.float32)
        elif len(shape) >= 2:  # conv / linear weight: fan-in scaled normal
            fan_in = int(np.prod(shape[1:]))
            out[key] = (rs.standard_normal(shape) * np.sqrt(1.5 / fan_in)).astype(np.float32)
        elif leaf == "weight":  # norm scale
            out[key] = rs.uniform(0.5, 1.5, size=shape).astype(np.float32)
        else:  # biases, norm shifts
            out[key] = (0.1 * rs.standard_normal(shape)).astype(np.float32)
    return out


# ------------------------------------------------------------------------------------------------
# state_dict layouts (key -> shape) of the reference's hot modules; used to build
# parameter sets on the GPU box, where the reference tree does not exist.
# ------------------------------------------------------------------------------------------------
def sa_layout(mlp, use_xyz=True, prefix=""):
    """PointnetSAModuleVotes.state_dict() (pointnet2_modules.py:51-54, pytorch_utils.py:12-36)."""
    spec = list(mlp)
    if use_xyz:
        spec[0] += 3
    sd = {}
    for i in range(len(spec) - 1):
        p = "%smlp_module.layer%d." % (prefix, i)
        sd[p + "conv.weight"] = (spec[i + 1], spec[i], 1, 1)
        for k in ("weight", "bias", "running_mean", "running_var"):
            sd[p + "normlayer.bn." + k] = (spec[i + 1],)
        sd[p + "normlayer.bn.num_batches_tracked"] = ()
    return sd


def transformer_layout(cls, dp, dm, prefix=""):
    """state_dict() of the transformer_block/variants.py classes."""
    sd = {}

    def lin(name, i, o, bias=True):
        sd[prefix + name + ".weight"] = (o, i)
        if bias:
            sd[prefix + name + ".bias"] = (o,)

    if cls == "TransformerBlockMLP":
        lin("fc1.0", dp, dm), lin("fc1.2", dm, dm), lin("fc2.0", dm, dm), lin("fc2.2", dm, dp)
    else:
        lin("fc1", dp, dm), lin("fc2", dm, dp)
    lin("fc_delta.0", 3, dm), lin("fc_delta.2", dm, dm)
    if cls != "TransformerBlockSTD":
        lin("fc_gamma.0", dm, dm), lin("fc_gamma.2", dm, dm)
    lin("w_qs", dm, dm, False), lin("w_ks", dm, dm, False), lin("w_vs", dm, dm, False)
    return sd


def hot_path_layout():
    """The hot-path part of a PTT tracker's state_dict under tools/cfgs/kitti_models/ptt.yaml."""
    sd = {}
    for l, mlp in enumerate(([0, 64, 64, 128], [128, 128, 128, 256], [256, 128, 128, 256])):
        sd.update(sa_layout(mlp, prefix="backbone_3d.SA_modules.%d." % l))
    sd["backbone_3d.cov_final.weight"] = (256, 256, 1)
    sd["backbone_3d.cov_final.bias"] = (256,)
    sd.update(transformer_layout("TransformerBlock", 256, 512, "centroid_voting_head.transformer_block."))
    sd.update(sa_layout([257, 256, 256, 256], prefix="box_voting_head.vote_aggregation."))
    sd.update(transformer_layout("TransformerBlock", 256, 512, "box_voting_head.transformer_block."))
    return sd


def seq_layout(channels, bn_last=False, prefix=""):
    """pytorch_utils.Seq of Conv1d layers: BatchNorm (no conv bias) on every layer but the last, which has a bias."""
    sd = {}
    n = len(channels) - 1
    for i in range(n):
        p = "%s%d." % (prefix, i)
        sd[p + "conv.weight"] = (channels[i + 1], channels[i], 1)
        if i < n - 1 or bn_last:
            for k in ("weight", "bias", "running_mean", "running_var"):
                sd[p + "normlayer.bn." + k] = (channels[i + 1],)
            sd[p + "normlayer.bn.num_batches_tracked"] = ()
        else:
            sd[p + "conv.bias"] = (channels[i + 1],)
    return sd


def full_model_layout():
    """Every parameter of a PTT tracker under tools/cfgs/kitti_models/ptt.yaml that the eval forward reads: the hot
    path plus the similarity module and the heads' Conv1d stacks."""
    sd = hot_path_layout()
    mlp = sa_layout([260, 256, 256, 256], use_xyz=False, prefix="similarity_module.")
    sd.update({k.replace("similarity_module.mlp_module.", "similarity_module.mlp."): v for k, v in mlp.items()})
    sd.update(seq_layout([256, 256, 256], prefix="similarity_module.conv."))
    sd.update(seq_layout([256, 256, 256, 1], prefix="centroid_voting_head.cla_layer."))
    sd.update(seq_layout([259, 256, 256, 259], prefix="centroid_voting_head.vote_layer."))
    sd.update(seq_layout([256, 256, 256, 5], prefix="box_voting_head.refine_layer."))
    return sd


def full_model_state_dict(seed=0):
    import torch

    return {k: torch.from_numpy(v) for k, v in fill_state_dict(full_model_layout(), seed).items()}


def hot_path_state_dict(seed=0):
    """Filled hot-path parameters as torch CPU tensors."""
    import torch

    return {k: torch.from_numpy(v) for k, v in fill_state_dict(hot_path_layout(), seed).items()}


def load_filled(module, seed=0):
    """Fill `module` in place from fill_state_dict (any torch.nn.Module)."""
    import torch

    sd = module.state_dict()
    filled = fill_state_dict(sd, seed)
    module.load_state_dict({k: torch.from_numpy(v).to(sd[k].device) for k, v in filled.items()})
    return module


def features(shape, seed=0, scale=1.0):
    """Seeded float32 normal tensor (numpy) used as per-point input features in tests."""
    rs = np.random.RandomState((seed * 7919 + 13) % (2 ** 31))
    return (scale * rs.standard_normal(shape)).astype(np.float32)


def crc(a):
    """Checksum of an array's bytes; fixtures carry it for regenerated inputs to detect recipe drift."""
    return zlib.crc32(np.ascontiguousarray(a).tobytes())
