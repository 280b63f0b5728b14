"""Loads the reference's own `ptt.models`.  TEST INFRASTRUCTURE.

Where the tree comes from: $PTT_REFERENCE_ROOT, else /root/reference (the authoring container), else oracle/_ref (the
copy oracle/make_ref.sh stages; git-ignored, it travels to the GPU box with the snapshot).  Used by
tests/golden/make_golden.py, by the tests marked `reference` and by bench.py's reference arm.

    load()                CPU: three imports missing from this image are shimmed (oracle/shims: thop, easydict,
                          pointnet2_ops._ext -> the CPU oracle) and `.cuda()` is neutralised (hard-coded at
                          pointnet2_modules.py:69,71, voting_head_template.py:23,25, ptt/models/__init__.py:17,19)
    load(device="cuda")   GPU box: thop / easydict shims only; `pointnet2_ops._ext` is the PRODUCT drop-in
                          (ptt_b200.install_dropin) and `.cuda()` is the real thing -- the reference's own modules then
                          run on the B200 over our C ABI exactly as they would over upstream pointnet2_ops
"""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIMS = os.path.join(_HERE, "shims")
_REPO = os.path.dirname(_HERE)


def _find_root():
    cands = [os.environ.get("PTT_REFERENCE_ROOT"), "/root/reference", os.path.join(_HERE, "_ref")]
    for c in cands:
        if c and os.path.isdir(os.path.join(c, "ptt", "models")):
            return c
    return cands[1]


REFERENCE_ROOT = _find_root()
_ORIG_CUDA = {}


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "ptt", "models"))


def _load_cuda():
    """The reference's modules on the GPU over the product drop-in (see the module docstring)."""
    import torch

    if getattr(torch.Tensor.cuda, "_oracle_identity", False):      # a CPU-mode load() ran earlier in this process
        torch.Tensor.cuda = _ORIG_CUDA["tensor"]
        torch.nn.Module.cuda = _ORIG_CUDA["module"]
    for p in (REFERENCE_ROOT, _SHIMS, _REPO):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    import ptt_b200

    ext = ptt_b200.install_dropin()           # puts ptt_b200/dropin ahead of oracle/shims for `pointnet2_ops`
    import ptt.models as models  # noqa: E402
    import ptt.models.backbones_3d.pointnet2.pointnet2_utils as pu  # noqa: E402

    pu._ext = ext
    return models


def load(device="cpu"):
    """Returns the imported `ptt.models` package of the reference."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    import torch

    if device != "cpu":
        return _load_cuda()
    for p in (REFERENCE_ROOT, _SHIMS, _REPO):
        if p not in sys.path:
            sys.path.insert(0, p)
    # shims must win over anything else called pointnet2_ops / thop / easydict
    sys.path.remove(_SHIMS)
    sys.path.insert(0, _SHIMS)
    if not getattr(torch.Tensor.cuda, "_oracle_identity", False):
        def _ident(self, *a, **k):
            return self
        _ident._oracle_identity = True
        _ORIG_CUDA["tensor"], _ORIG_CUDA["module"] = torch.Tensor.cuda, torch.nn.Module.cuda
        torch.Tensor.cuda = _ident
        torch.nn.Module.cuda = _ident
    # another `pointnet2_ops` (e.g. the product drop-in, installed by a test in the same process) must not win here
    for name in [m for m in sys.modules if m == "pointnet2_ops" or m.startswith("pointnet2_ops.")]:
        if not getattr(sys.modules[name], "__file__", "").startswith(_SHIMS):
            del sys.modules[name]
    import pointnet2_ops._ext as shim_ext  # noqa: E402
    import ptt.models as models  # noqa: E402
    import ptt.models.backbones_3d.pointnet2.pointnet2_utils as pu  # noqa: E402

    pu._ext = shim_ext
    return models


def load_cfg(name="kitti_models/ptt.yaml"):
    """Fresh EasyDict of a reference YAML (fresh per model build: PointnetSAModuleVotes.__init__
    mutates its mlp list in place, pointnet2_modules.py:51-53).  Call load() first."""
    import yaml
    from easydict import EasyDict

    with open(os.path.join(REFERENCE_ROOT, "tools", "cfgs", name)) as f:
        return EasyDict(yaml.safe_load(f))


class _DatasetStandIn:
    """The attributes Tracker3DTemplate reads (tracker3d_template.py:14-16,34-41)."""

    training = False
    class_names = ["Car"]
    grid_size = None
    voxel_size = None
    point_cloud_range = None

    class point_feature_encoder:
        num_point_features = 3


def build_tracker(training=False, device="cpu"):
    """The reference's own PTT tracker from tools/cfgs/kitti_models/ptt.yaml through its own factory
    (ptt/models/__init__.py:9-10); on `device`."""
    models = load(device)
    cfg = load_cfg()
    ds = _DatasetStandIn()
    ds.training = training
    net = models.build_network(cfg.MODEL, 1, ds)
    net.train(training)
    if device != "cpu":
        net = net.to(device)
    return net, cfg
