"""The hot path driven through the REFERENCE's own nn.Modules (oracle/refload.py).  TEST INFRASTRUCTURE.

`RefHotPath` builds the reference's PTT tracker with its own factory and runs, with the reference's own module
objects, exactly the stages ptt_b200.hotpath.HotPath.forward runs (same glue between them as
tests/golden/make_golden.py::hot_path_fixture), or the whole tracker forward:

    device="cpu"    the reference's modules on the host cores; its CUDA-only pointnet2_ops replaced by the oracle's C
                    ops (oracle/pointnet2_ref.c) -- bench.py --impl reference / cpu_baseline
    device="cuda"   the reference's modules on the B200 (cuDNN / cuBLAS / eager torch) over the product
                    `pointnet2_ops._ext` drop-in -- the honest "unfused reference on the same GPU" figure, and the
                    Level-1 drop-in parity tests
"""
import torch

from . import refload


class RefHotPath:
    def __init__(self, state_dict=None, device="cpu", register_b200=False):
        """state_dict: tensors keyed like the reference model (missing keys keep their initial values).
        register_b200: install ptt_b200's modules into the reference's registries first (Level 2 drop-in)."""
        self.device = torch.device(device)
        kind = "cpu" if self.device.type == "cpu" else str(self.device)
        refload.load(kind)
        self._restore = None
        if register_b200:
            from ptt.models import transformer_block
            from ptt.models.backbones_3d.pointnet2 import pointnet2_modules
            from ptt_b200 import modules as m

            self._restore = (pointnet2_modules.PointnetSAModuleVotes, dict(transformer_block.__all__))
            m.register(install_ext=False)
        try:
            self.net, self.cfg = refload.build_tracker(training=False, device=kind)
        finally:
            if self._restore is not None:       # leave the reference's registries as we found them
                from ptt.models import transformer_block
                from ptt.models.backbones_3d.pointnet2 import pointnet2_modules

                pointnet2_modules.PointnetSAModuleVotes = self._restore[0]
                transformer_block.__all__.clear()
                transformer_block.__all__.update(self._restore[1])
        if state_dict is not None:
            own = self.net.state_dict()
            sd = {k: v.to(own[k].device, own[k].dtype) for k, v in state_dict.items() if k in own}
            self.net.load_state_dict(sd, strict=False)
        self.net.eval()

    @torch.no_grad()
    def hot_path(self, search, template):
        """search (B,Ns,3), template (B,Nt,3) on self.device -> the dict HotPath.forward returns."""
        net, cfg = self.net, self.cfg
        bd = net.backbone_3d({"search_points": search, "template_points": template})
        cen = net.centroid_voting_head.transformer_block(
            xyz=bd["search_seeds"], features=bd["search_feats"].transpose(1, 2).contiguous())[0]
        votes_feats = torch.cat([torch.full_like(cen[:, :, :1], 0.5), cen], dim=2).transpose(1, 2).contiguous()
        b_xyz, b_feat, _ = net.box_voting_head.vote_aggregation(
            xyz=bd["search_seeds"], features=votes_feats, npoint=cfg.MODEL.BOX_HEAD.SA_CONFIG.NPOINTS)
        box = net.box_voting_head.transformer_block(xyz=b_xyz, features=b_feat.transpose(1, 2).contiguous())[0]
        out = {k: bd[k] for k in ("search_seeds", "search_feats", "search_inds", "template_seeds", "template_feats",
                                  "template_inds")}
        out.update(centroid_feats=cen, box_centers=b_xyz, box_sa_feats=b_feat, box_feats=box)
        return out

    @torch.no_grad()
    def full(self, search, template):
        """The reference's own PTT.forward in eval mode (trackers/ptt.py:42-51) -> its batch_dict."""
        return self.net({"search_points": search, "template_points": template})
