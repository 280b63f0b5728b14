"""CPU restatement ("port") of the Python half of the PTT hot path.  TEST INFRASTRUCTURE.

The reference is Python and cannot travel to the GPU box, so the checker there is this port:
plain PyTorch CPU code over the C ops of oracle/pointnet2_ref.c, written as functions of a
state_dict whose keys are the reference's own (so a reference checkpoint drives it unchanged).
It is PINNED in this container against the reference's modules imported from /root/reference
(tests/test_oracle_vs_reference.py) and against the committed golden fixtures the reference
produced (tests/golden/*.npz, made by tests/golden/make_golden.py).

Citations are file:line under /root/reference/ptt/models/.
"""
import math

import torch
import torch.nn.functional as F

from . import cops


def _sub(sd, prefix):
    n = len(prefix)
    return {k[n:]: v for k, v in sd.items() if k.startswith(prefix)}


# ----------------------------------------------------------------------------------------------
# a5  QueryAndGroup.forward            backbones_3d/pointnet2/pointnet2_utils.py:320-380
# ----------------------------------------------------------------------------------------------
def query_and_group(xyz, new_xyz, features, radius, nsample, use_xyz=True, normalize_xyz=False):
    """xyz (B,N,3), new_xyz (B,M,3), features (B,C,N)|None -> new_features (B,3+C,M,ns), grouped_xyz, idx."""
    idx = cops.ball_query(new_xyz.contiguous(), xyz.contiguous(), float(radius), int(nsample))  # :337
    grouped_xyz = cops.group_points(xyz.transpose(1, 2).contiguous(), idx)                     # :350-351
    grouped_xyz = grouped_xyz - new_xyz.transpose(1, 2).unsqueeze(-1)                           # :352
    if normalize_xyz:
        grouped_xyz = grouped_xyz / radius                                                       # :353-354
    if features is not None:
        grouped = cops.group_points(features.contiguous(), idx)                                 # :357
        new_features = torch.cat([grouped_xyz, grouped], dim=1) if use_xyz else grouped         # :358-363
    else:
        new_features = grouped_xyz                                                              # :364-368
    return new_features, grouped_xyz, idx


# ----------------------------------------------------------------------------------------------
# a7  SharedMLP = [1x1 Conv2d(no bias) -> BatchNorm2d -> ReLU] * L    pytorch_utils.py:12-36,158-189
# ----------------------------------------------------------------------------------------------
def shared_mlp(sd, x, training=False, momentum=0.1, eps=1e-5):
    """sd keys: layer{i}.conv.weight, layer{i}.normlayer.bn.{weight,bias,running_mean,running_var}."""
    i = 0
    while "layer%d.conv.weight" % i in sd:
        p = "layer%d." % i
        x = F.conv2d(x, sd[p + "conv.weight"], sd.get(p + "conv.bias"))
        if p + "normlayer.bn.weight" in sd:
            x = F.batch_norm(x, sd[p + "normlayer.bn.running_mean"], sd[p + "normlayer.bn.running_var"],
                             sd[p + "normlayer.bn.weight"], sd[p + "normlayer.bn.bias"],
                             training=training, momentum=momentum, eps=eps)
        x = F.relu(x)
        i += 1
    return x


# ----------------------------------------------------------------------------------------------
# a6  PointnetSAModuleVotes.forward    backbones_3d/pointnet2/pointnet2_modules.py:57-90
# ----------------------------------------------------------------------------------------------
def sa_module_votes(sd, xyz, features, npoint, radius, nsample, sample_method="fps", use_xyz=True,
                    normalize_xyz=False, inds=None, training=False):
    """-> new_xyz (B,npoint,3), new_features (B,Cout,npoint), inds int64 (B,npoint)."""
    B = xyz.shape[0]
    if inds is None:
        if sample_method == "fps":
            inds = cops.furthest_point_sampling(xyz.contiguous(), npoint)                       # :72-73
        elif sample_method in ("sequence", "rs"):
            inds = torch.arange(npoint, dtype=torch.int32).repeat(B, 1)                          # :68-71
        else:
            raise NotImplementedError(sample_method)
    new_xyz = cops.gather_points(xyz.transpose(1, 2).contiguous(), inds.int().contiguous())     # :79
    new_xyz = new_xyz.transpose(1, 2).contiguous()                                              # :81
    grouped, _, _ = query_and_group(xyz, new_xyz, features, radius, nsample, use_xyz, normalize_xyz)  # :83
    y = shared_mlp(_sub(sd, "mlp_module."), grouped, training=training)                         # :84
    y = y.max(dim=3)[0]                                                                          # :85-88
    return new_xyz, y, inds.to(torch.int64)                                                      # :90


# ----------------------------------------------------------------------------------------------
# a8  PointNet2BackboneLight.branch_forward    backbones_3d/pointnet2_backbone.py:41-50
# ----------------------------------------------------------------------------------------------
def backbone_branch(sd, pts, npoints, radii=(0.3, 0.5, 0.7), nsamples=(32, 32, 32),
                    sample_methods=("fps", "sequence", "sequence"), normalize_xyz=True, training=False):
    """pts (B,N,3) -> seeds (B,n3,3), point_features (B,256,n3), inds int64 (B,n3)."""
    xyz, feats = pts[..., 0:3].contiguous(), None
    inds = []
    for l in range(3):
        xyz, feats, i = sa_module_votes(_sub(sd, "SA_modules.%d." % l), xyz, feats, npoints[l], radii[l],
                                        nsamples[l], sample_methods[l], True, normalize_xyz,
                                        training=training)
        inds.append(i)
    point_features = F.conv1d(feats, sd["cov_final.weight"], sd["cov_final.bias"])              # :46
    composed = inds[0].gather(1, inds[1]).gather(1, inds[2])                                    # :48
    return xyz, point_features, composed


# ----------------------------------------------------------------------------------------------
# a9  TransformerBlock.forward (kNN vector attention)    transformer_block/variants.py:127-165
# ----------------------------------------------------------------------------------------------
def _gather_rows(points, idx):
    """index_points (model_utils/layer_utils.py:29-40): points (B,N,C), idx (B,S,K) -> (B,S,K,C)."""
    B, S, K = idx.shape
    flat = idx.reshape(B, S * K, 1).expand(-1, -1, points.shape[-1])
    return torch.gather(points, 1, flat).reshape(B, S, K, -1)


def _linear(sd, name, x):
    return F.linear(x, sd[name + ".weight"], sd.get(name + ".bias"))


def _mlp2(sd, name, x):
    return _linear(sd, name + ".2", F.relu(_linear(sd, name + ".0", x)))


def knn_indices(xyz, k):
    """argsort of square_distance (layer_utils.py:12-26, variants.py:150-151) with ties resolved
    lowest-index-first (the reference's argsort is unstable, so its tie order is unspecified)."""
    return cops.knn(xyz.contiguous(), int(k)).to(torch.int64)


def transformer_block(sd, xyz, features, k, variant="TransformerBlock", knn_idx=None):
    """xyz (B,n,3), features (B,n,d_points) -> (res (B,n,d_points), attn (B,n,k,d_model)).

    variant: 'TransformerBlock' (variants.py:127-165), 'TransformerBlockMLP' (:211-256, two-layer
    fc1/fc2), 'TransformerBlockOffset' (:297-334, fc2(x - res))."""
    if knn_idx is None:
        knn_idx = knn_indices(xyz, k)                                                            # :150-151
    knn_xyz = _gather_rows(xyz, knn_idx)                                                          # :152
    pre = features
    x = _mlp2(sd, "fc1", features) if variant == "TransformerBlockMLP" else _linear(sd, "fc1", features)  # :155
    q = _linear(sd, "w_qs", x)
    kk = _gather_rows(_linear(sd, "w_ks", x), knn_idx)
    v = _gather_rows(_linear(sd, "w_vs", x), knn_idx)                                             # :156
    pos = _mlp2(sd, "fc_delta", xyz[:, :, None] - knn_xyz)                                        # :158
    attn = _mlp2(sd, "fc_gamma", q[:, :, None] - kk + pos)                                        # :160
    attn = F.softmax(attn / math.sqrt(kk.size(-1)), dim=-2)                                       # :161
    res = torch.einsum("bmnf,bmnf->bmf", attn, v + pos)                                           # :163
    if variant == "TransformerBlockOffset":
        res = x - res
    res = (_mlp2(sd, "fc2", res) if variant == "TransformerBlockMLP" else _linear(sd, "fc2", res)) + pre  # :164
    return res, attn


# a10 TransformerBlockSTD.forward (dense n x n dot-product attention)    variants.py:12-40
def transformer_block_std(sd, xyz, features):
    pre = features
    x = _linear(sd, "fc1", features)
    q, k, v = _linear(sd, "w_qs", x), _linear(sd, "w_ks", x), _linear(sd, "w_vs", x)
    attn = F.softmax(q @ k.transpose(1, 2) / math.sqrt(k.size(-1)), dim=-1)
    res = attn @ (v + _mlp2(sd, "fc_delta", xyz))
    return _linear(sd, "fc2", res) + pre, attn


# ----------------------------------------------------------------------------------------------
# The bench / smoke "frame": the hot-path functions a1-a9 in dependency order, with the non-hot
# modules between them (CosineSimAug, the heads' Conv1d stacks -- SURVEY.md 8(f) N1/N2) replaced by
# fixed cheap glue so that real tensors flow from one hot stage into the next:
#   search/template clouds -> backbone (SA1-3 x 2 branches)                    pointnet2_backbone.py:52-67
#   centroid-head transformer on (search_seeds, search_feats^T)                centroids_voting_head.py:71-76
#   box-head SA on (votes = search_seeds, votes_feats = [score=0.5 | feats])   box_voting_head.py:75-79
#   box-head transformer on (centres, proposal_feats^T)                        box_voting_head.py:81-86
# ptt_b200.hotpath.HotPath.forward is the product twin of this function.
# ----------------------------------------------------------------------------------------------
def hot_path_frame(sd, search, template, cfg=None):
    c = dict(npoints_search=(512, 256, 128), npoints_template=(256, 128, 64), radii=(0.3, 0.5, 0.7),
             nsamples=(32, 32, 32), knn=16, box_npoint=64, box_radius=0.3, box_nsample=16)
    c.update(cfg or {})
    bb = _sub(sd, "backbone_3d.")
    s_xyz, s_feat, s_inds = backbone_branch(bb, search, c["npoints_search"], c["radii"], c["nsamples"])
    t_xyz, t_feat, t_inds = backbone_branch(bb, template, c["npoints_template"], c["radii"], c["nsamples"])
    cen, _ = transformer_block(_sub(sd, "centroid_voting_head.transformer_block."), s_xyz,
                               s_feat.transpose(1, 2).contiguous(), c["knn"])
    votes_feats = torch.cat([torch.full_like(cen[:, :, :1], 0.5), cen], dim=2).transpose(1, 2).contiguous()
    b_xyz, b_feat, _ = sa_module_votes(_sub(sd, "box_voting_head.vote_aggregation."), s_xyz, votes_feats,
                                       c["box_npoint"], c["box_radius"], c["box_nsample"], "fps", True, True)
    box, _ = transformer_block(_sub(sd, "box_voting_head.transformer_block."), b_xyz,
                               b_feat.transpose(1, 2).contiguous(), c["knn"])
    return {"search_seeds": s_xyz, "search_feats": s_feat, "search_inds": s_inds,
            "template_seeds": t_xyz, "template_feats": t_feat, "template_inds": t_inds,
            "centroid_feats": cen, "box_centers": b_xyz, "box_sa_feats": b_feat, "box_feats": box}


# ----------------------------------------------------------------------------------------------
# the modules between the hot stages, so that the checker covers the whole tracker forward.
# ----------------------------------------------------------------------------------------------
def seq_conv1d(sd, x, training=False, eps=1e-5):
    """pytorch_utils.Seq of Conv1d layers (pytorch_utils.py:270-300; keys {i}.conv.weight (Cout,Cin,1),
    {i}.conv.bias, {i}.normlayer.bn.*).  ReLU after every layer but the last (the heads and the similarity
    module all end with activation=None: centroids_voting_head.py:14-25, box_voting_head.py:24-29,
    p2b_xcoor.py:20-24).  x (B,C,N)."""
    n = 0
    while "%d.conv.weight" % n in sd:
        n += 1
    for i in range(n):
        p = "%d." % i
        x = F.conv1d(x, sd[p + "conv.weight"], sd.get(p + "conv.bias"))
        if p + "normlayer.bn.weight" in sd:
            x = F.batch_norm(x, sd[p + "normlayer.bn.running_mean"], sd[p + "normlayer.bn.running_var"],
                             sd[p + "normlayer.bn.weight"], sd[p + "normlayer.bn.bias"], training=training, eps=eps)
        if i < n - 1:
            x = F.relu(x)
    return x


def cosine_sim_aug(sd, search_feats, template_feats, template_xyz):
    """CosineSimAug.forward (similarity_modules/p2b_xcoor.py:25-46).  search_feats (B,f,n2), template_feats (B,f,n1),
    template_xyz (B,n1,3) -> cosine_feats (B,256,n2)."""
    b, f, n2 = search_feats.shape
    n1 = template_feats.shape[-1]
    sim = F.cosine_similarity(template_feats.unsqueeze(-1).expand(b, f, n1, n2),
                              search_feats.unsqueeze(2).expand(b, f, n1, n2), dim=1)                # :36-37
    xyz_ = template_xyz.transpose(1, 2).contiguous().unsqueeze(-1).expand(b, 3, n1, n2)             # :38
    fusion = torch.cat((sim.unsqueeze(1), xyz_), dim=1)                                             # :39
    fusion = torch.cat((fusion, template_feats.unsqueeze(-1).expand(b, f, n1, n2)), dim=1)          # :40
    fusion = shared_mlp(_sub(sd, "mlp."), fusion)                                                    # :41
    fusion = fusion.max(dim=2)[0]                                                                    # :42-43
    return seq_conv1d(_sub(sd, "conv."), fusion)                                                     # :44


def centroid_voting_head(sd, search_seeds, cosine_feats, k):
    """CentroidVotingHead.forward (voting_heads/centroids_voting_head.py:66-100), CLS_USE_SEARCH_XYZ False.
    -> pred_centroids_cls (B,n), pred_centroids_votes (B,n,3), votes_feats (B,257,n), trans_feat (B,n,256)."""
    xyz_cm = search_seeds.transpose(1, 2).contiguous()
    trans, _ = transformer_block(_sub(sd, "transformer_block."), search_seeds, cosine_feats.transpose(1, 2).contiguous(), k)
    fusion = trans.transpose(1, 2).contiguous()
    cls_out = seq_conv1d(_sub(sd, "cla_layer."), fusion).squeeze(1)                                  # :87
    score = cls_out.sigmoid()
    voting_input = torch.cat((xyz_cm, fusion), dim=1)                                                # :92
    voting_results = voting_input + seq_conv1d(_sub(sd, "vote_layer."), voting_input)               # :93-95
    votes = voting_results[:, 0:3, :].transpose(1, 2).contiguous()
    votes_feats = torch.cat((score.unsqueeze(1), voting_results[:, 3:, :]), dim=1)                   # :100
    return cls_out, votes, votes_feats, trans


def box_voting_head(sd, votes, votes_feats, c):
    """BoxVotingHead.forward, eval (voting_heads/box_voting_head.py:70-95)."""
    b_xyz, b_feat, _ = sa_module_votes(_sub(sd, "vote_aggregation."), votes, votes_feats, c["box_npoint"],
                                       c["box_radius"], c["box_nsample"], "fps", True, True)
    box, _ = transformer_block(_sub(sd, "transformer_block."), b_xyz, b_feat.transpose(1, 2).contiguous(), c["knn"])
    off = seq_conv1d(_sub(sd, "refine_layer."), box.transpose(1, 2).contiguous())                    # :88
    est = torch.cat((off[:, 0:3, :] + b_xyz.transpose(1, 2).contiguous(), off[:, 3:, :]), dim=1)     # :90-91
    return b_xyz, est.transpose(1, 2).contiguous(), b_feat, box


def full_model_frame(sd, search, template, cfg=None):
    """The whole PTT tracker forward in eval mode (trackers/ptt.py:45-46 over the module list of ptt.yaml):
    backbone -> CosineSimAug -> CentroidVotingHead -> BoxVotingHead.  ptt_b200.hotpath.HotPath.forward_full is the
    product twin."""
    c = dict(npoints_search=(512, 256, 128), npoints_template=(256, 128, 64), radii=(0.3, 0.5, 0.7),
             nsamples=(32, 32, 32), knn=16, box_npoint=64, box_radius=0.3, box_nsample=16)
    c.update(cfg or {})
    bb = _sub(sd, "backbone_3d.")
    s_xyz, s_feat, s_inds = backbone_branch(bb, search, c["npoints_search"], c["radii"], c["nsamples"])
    t_xyz, t_feat, t_inds = backbone_branch(bb, template, c["npoints_template"], c["radii"], c["nsamples"])
    cos = cosine_sim_aug(_sub(sd, "similarity_module."), s_feat, t_feat, t_xyz)
    cls_out, votes, votes_feats, cen = centroid_voting_head(_sub(sd, "centroid_voting_head."), s_xyz, cos, c["knn"])
    b_xyz, box_data, b_feat, box = box_voting_head(_sub(sd, "box_voting_head."), votes, votes_feats, c)
    return {"search_seeds": s_xyz, "search_feats": s_feat, "search_inds": s_inds,
            "template_seeds": t_xyz, "template_feats": t_feat, "template_inds": t_inds,
            "cosine_feats": cos, "centroid_feats": cen, "pred_centroids_cls": cls_out, "pred_centroids_votes": votes,
            "votes_feats": votes_feats, "pred_box_center": b_xyz, "box_sa_feats": b_feat, "box_feats": box,
            "pred_box_data": box_data}


# ----------------------------------------------------------------------------------------------
# a10 / N4: the other blocks of transformer_block/__init__.py:7-17
# ----------------------------------------------------------------------------------------------
def transformer_block_cosine(sd, xyz, features, k):
    """TransformerBlockCosine.forward (variants.py:66-88)."""
    knn_idx = knn_indices(xyz, k)
    x = _linear(sd, "fc1", features)
    q, kk, v = _linear(sd, "w_qs", x), _gather_rows(_linear(sd, "w_ks", x), knn_idx), _gather_rows(_linear(sd, "w_vs", x), knn_idx)
    pos = _mlp2(sd, "fc_delta", xyz[:, :, None] - _gather_rows(xyz, knn_idx))
    sim = F.cosine_similarity(q.unsqueeze(-2).repeat(1, 1, k, 1), kk, dim=-1)                      # :78-79
    rel = _linear(sd, "fc_sim", torch.cat((sim.unsqueeze(-1), q[:, :, None] - kk), dim=-1))         # :80-82
    attn = F.softmax(_mlp2(sd, "fc_gamma", rel + pos) / math.sqrt(kk.size(-1)), dim=-2)             # :83-84
    res = torch.einsum("bmnf,bmnf->bmf", attn, v + pos)
    return _linear(sd, "fc2", res) + features, attn


def transformer_block_all(sd, xyz, features):
    """TransformerBlockALL.forward (variants.py:111-124): softmax over the n tokens (dim=-2 of (b, n, d_model))."""
    x = _linear(sd, "fc1", features)
    q, kk, v = _linear(sd, "w_qs", x), _linear(sd, "w_ks", x), _linear(sd, "w_vs", x)
    pos = _mlp2(sd, "fc_delta", xyz)
    attn = F.softmax(_mlp2(sd, "fc_gamma", q - kk + pos) / math.sqrt(kk.size(-1)), dim=-2)
    return _linear(sd, "fc2", attn * (v + pos)) + features, attn


def cross_attention_block(sd, xyz, search_feat, template_feat, k):
    """CrossAttentionBlock.forward (variants.py:190-208)."""
    knn_idx = knn_indices(xyz, k)
    s, tm = _linear(sd, "fc1", search_feat), _linear(sd, "fc1", template_feat)
    q = _linear(sd, "w_qs", tm)
    kk, v = _gather_rows(_linear(sd, "w_ks", s), knn_idx), _gather_rows(_linear(sd, "w_vs", s), knn_idx)
    pos = _mlp2(sd, "fc_delta", xyz[:, :, None] - _gather_rows(xyz, knn_idx))
    attn = F.softmax(_mlp2(sd, "fc_gamma", q[:, :, None] - kk + pos) / math.sqrt(kk.size(-1)), dim=-2)
    res = torch.einsum("bmnf,bmnf->bmf", attn, v + pos)
    return _linear(sd, "fc3", res) + search_feat, attn


def transformer_block_backbone(sd, new_xyz, grouped_xyz, grouped_idx, features):
    """TransformerBlockBackbone.forward (variants.py:281-294), without its debug prints."""
    idx = grouped_idx.long()
    x = _linear(sd, "fc1", features)
    q, kk, v = _linear(sd, "w_qs", x), _gather_rows(_linear(sd, "w_ks", x), idx), _gather_rows(_linear(sd, "w_vs", x), idx)
    pos = _mlp2(sd, "fc_delta", new_xyz[:, :, None] - grouped_xyz.permute(0, 2, 3, 1).contiguous())
    attn = F.softmax(_mlp2(sd, "fc_gamma", q[:, :, None] - kk + pos) / math.sqrt(kk.size(-1)), dim=-2)
    return torch.einsum("bmnf,bmnf->bmf", attn, v + pos).contiguous()


def mul_head_transformer_layer(sd, xyz, features, k, heads, eps=1e-5):
    """MulHeadTransformerLayer.forward (multitransformer.py:38-63), dropout p = 0."""
    knn_idx = knn_indices(xyz, k)
    x = _linear(sd, "fc1", features)
    B, N, C = x.shape
    query = _linear(sd, "w_qs", x).view(B, N, heads, -1).permute(0, 2, 1, 3).flatten(0, 1)

    def split(t):
        return t.view(B, N, t.shape[2], heads, -1).permute(0, 3, 1, 2, 4).flatten(0, 1)                # :51-53

    key, value = split(_gather_rows(_linear(sd, "w_ks", x), knn_idx)), split(_gather_rows(_linear(sd, "w_vs", x), knn_idx))
    pos = split(_mlp2(sd, "fc_delta", xyz[:, :, None] - _gather_rows(xyz, knn_idx)))
    attn = F.softmax(_mlp2(sd, "fc_gamma", query[:, :, None] - key + pos) / math.sqrt(key.size(-1)), dim=-2)
    res = torch.einsum("bmnf,bmnf->bmf", attn, value + pos)
    if heads > 1:
        res = res.permute(0, 2, 1).reshape(B, C, N).permute(0, 2, 1)                                    # :59-60
    res = F.layer_norm(F.linear(res, sd["proj.weight"]), (C,), sd["norm1.weight"], sd["norm1.bias"], eps)
    dp = features.shape[2]
    return F.layer_norm(_linear(sd, "fc2", res), (dp,), sd["norm2.weight"], sd["norm2.bias"], eps) + features, attn


def mul_transformer_block(sd, xyz, features, k, heads):
    """MulTransformerBlock.forward (multitransformer.py:72-76); keys layers.{i}.*."""
    out, attn, i = features, None, 0
    while "layers.%d.fc1.weight" % i in sd:
        out, attn = mul_head_transformer_layer(_sub(sd, "layers.%d." % i), xyz, out, k, heads)
        i += 1
    return out, attn
