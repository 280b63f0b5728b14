"""Authoring container only: oracle/torch_port.py against the reference's OWN modules imported live from
/root/reference (fresh seeded inputs, beyond the committed fixtures)."""
import numpy as np
import pytest
import torch

from conftest import t
from oracle import refload, torch_port
from ptt_b200 import synth

pytestmark = pytest.mark.reference
TOL = dict(rtol=1e-4, atol=1e-4)


def test_port_matches_live_reference_tracker_hot_path():
    torch.set_grad_enabled(False)
    net, cfg = refload.build_tracker(training=False)
    synth.load_filled(net, seed=5)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    for kind, ns, nt, seed in (("dense", 1024, 512, 80), ("sparse", 512, 512, 81)):
        search = t(synth.make_clouds(1, ns, seed, kind))
        template = t(synth.make_clouds(1, nt, seed + 1, kind, role="template"))
        bd = net.backbone_3d({"search_points": search.clone(), "template_points": template.clone()})
        cen = net.centroid_voting_head.transformer_block(xyz=bd["search_seeds"],
                                                         features=bd["search_feats"].transpose(1, 2).contiguous())[0]
        got = torch_port.hot_path_frame(sd, search, template)
        assert torch.equal(got["search_inds"], bd["search_inds"]) and torch.equal(got["template_inds"], bd["template_inds"])
        assert torch.equal(got["search_seeds"], bd["search_seeds"])
        np.testing.assert_allclose(got["search_feats"].numpy(), bd["search_feats"].numpy(), **TOL)
        np.testing.assert_allclose(got["template_feats"].numpy(), bd["template_feats"].numpy(), **TOL)
        np.testing.assert_allclose(got["centroid_feats"].numpy(), cen.numpy(), **TOL)
    torch.set_grad_enabled(True)
