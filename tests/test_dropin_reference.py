"""Needs the reference tree (/root/reference, or the copy oracle/make_ref.sh stages): the drop-in seam really is where the reference looks.
No GPU here, so this checks resolution / registration / state_dict compatibility, not compute."""
import os
import subprocess
import sys

import pytest

from conftest import REPO
from oracle import refload

SCRIPT = r'''
import os, sys
repo = sys.argv[1]
ref = sys.argv[2]
sys.path[:0] = [repo, os.path.join(repo, "oracle", "shims"), ref]   # thop / easydict shims only
import torch
torch.Tensor.cuda = lambda self, *a, **k: self
torch.nn.Module.cuda = lambda self, *a, **k: self
import ptt_b200, ptt_b200.modules as m
ext = ptt_b200.install_dropin()
import ptt.models.backbones_3d.pointnet2.pointnet2_utils as pu          # the reference's own wrapper module
assert pu._ext is ext and ext.__file__.startswith(ptt_b200.DROPIN_DIR), pu._ext.__file__
ref_sa = __import__("ptt.models.backbones_3d.pointnet2.pointnet2_modules", fromlist=["x"]).PointnetSAModuleVotes
from ptt.models import transformer_block
ref_tr = transformer_block.__all__["TransformerBlock"]
pm, tb = m.register()
assert pm.PointnetSAModuleVotes is m.PointnetSAModuleVotes and tb.__all__["TransformerBlock"] is m.TransformerBlock
# a whole tracker built by the reference's own factory now contains our modules, with the reference's state_dict
import yaml
from easydict import EasyDict
cfg = EasyDict(yaml.safe_load(open(os.path.join(ref, "tools/cfgs/kitti_models/ptt.yaml"))))
class DS:
    training = False; class_names = ["Car"]; grid_size = voxel_size = point_cloud_range = None
    class point_feature_encoder: num_point_features = 3
from ptt.models import build_network
net = build_network(cfg.MODEL, 1, DS())
assert isinstance(net.backbone_3d.SA_modules[0], m.PointnetSAModuleVotes)
assert isinstance(net.centroid_voting_head.transformer_block, m.TransformerBlock)
assert isinstance(net.box_voting_head.vote_aggregation, m.PointnetSAModuleVotes)
ours = {k: tuple(v.shape) for k, v in net.state_dict().items()}
# the same model with the reference's own classes
pm.PointnetSAModuleVotes = ref_sa; tb.__all__["TransformerBlock"] = ref_tr
cfg = EasyDict(yaml.safe_load(open(os.path.join(ref, "tools/cfgs/kitti_models/ptt.yaml"))))
theirs = {k: tuple(v.shape) for k, v in build_network(cfg.MODEL, 1, DS()).state_dict().items()}
assert ours == theirs, set(ours) ^ set(theirs)
print("OK", len(ours), sum(1 for _ in net.parameters()))
'''


@pytest.mark.reference
def test_reference_resolves_ext_and_registries_to_ptt_b200():
    r = subprocess.run([sys.executable, "-c", SCRIPT, REPO, refload.REFERENCE_ROOT], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.strip().startswith("OK 175")          # 175 state_dict keys


REGISTRY_SCRIPT = r'''
import os, sys
repo = sys.argv[1]
ref = sys.argv[2]
sys.path[:0] = [repo, os.path.join(repo, "oracle", "shims"), ref]
import torch
torch.Tensor.cuda = lambda self, *a, **k: self
torch.nn.Module.cuda = lambda self, *a, **k: self
import ptt_b200.modules as m
from ptt.models import transformer_block
from easydict import EasyDict
ref = dict(transformer_block.__all__)                 # the reference's own classes, before registration
assert set(ref) == set(m.REGISTRY), set(ref) ^ set(m.REGISTRY)      # every registered name has a B200 twin
m.register()
n = 0
for name in sorted(ref):
    cfg = EasyDict(NAME=name, DIM_INPUT=32, DIM_MODEL=64, KNN=8, N_HEADS=4, N_LAYERS=2)
    ours = transformer_block.build_transformer(cfg)   # the reference's factory (transformer_block/__init__.py:20-27)
    assert type(ours) is m.REGISTRY[name], name
    theirs = ref[name](d_points=32, d_model=64, k=8, heads=4, layers=2)
    a = {k: tuple(v.shape) for k, v in ours.state_dict().items()}
    b = {k: tuple(v.shape) for k, v in theirs.state_dict().items()}
    assert a == b, (name, set(a) ^ set(b))            # checkpoints load by key and shape (tracker3d_template.py:110-118)
    n += 1
print("OK", n)
'''


@pytest.mark.reference
def test_every_registered_transformer_block_has_a_state_dict_compatible_twin():
    r = subprocess.run([sys.executable, "-c", REGISTRY_SCRIPT, REPO, refload.REFERENCE_ROOT], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.strip() == "OK 9"
