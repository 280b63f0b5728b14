"""Loads the reference's own `ptt.models` from /root/reference on CPU.  TEST INFRASTRUCTURE.

Only usable in the authoring container (/root/reference does not exist on the GPU box); used by
tests/golden/make_golden.py and by the container-only tests that pin oracle/torch_port.py.
Three imports missing from this image are shimmed (oracle/shims: thop, easydict,
pointnet2_ops._ext -> CPU oracle) and `.cuda()` is neutralised (hard-coded at
pointnet2_modules.py:69,71, voting_head_template.py:23,25, ptt/models/__init__.py:17,19).
"""
import os
import sys

REFERENCE_ROOT = os.environ.get("PTT_REFERENCE_ROOT", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")
_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "ptt", "models"))


def load():
    """Returns the imported `ptt.models` package of the reference."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    import torch

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
    mutates its mlp list in place, pointnet2_modules.py:51-53)."""
    import yaml
    from easydict import EasyDict

    load()
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


def build_tracker(training=False):
    models = load()
    cfg = load_cfg()
    ds = _DatasetStandIn()
    ds.training = training
    net = models.build_network(cfg.MODEL, 1, ds)
    net.train(training)
    return net, cfg
