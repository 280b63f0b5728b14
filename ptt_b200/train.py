"""Trainable twin of the hot path (the training row): the reference's module tree for rows a1-a9 --
backbone SA1-3 + cov_final (pointnet2_backbone.py:14-50), the two transformer blocks and the box-head SA
(centroids_voting_head.py:27-28, box_voting_head.py:15-31) -- assembled from ptt_b200.modules with the reference's
parameter names, so that `state_dict()` keys are those HotPath consumes and a reference checkpoint's hot-path entries
load by key.  In train() mode the SA layers and transformer blocks run the NATIVE training Functions of
ptt_b200/train_ops.py (forward and backward are library kernels; BatchNorm batch statistics and gradients as in the
reference; `module.native_train = False` selects the reference's decomposition over our ops under autograd, the
cross-check of the tests); under DistributedDataParallel the only collective of a step is the gradient all-reduce over
NCCL.  `time_train_step` is the step bench.py reports: by default captured once into a CUDA graph and replayed.
"""
import torch
import torch.nn as nn

from . import modules as m


class _Backbone(nn.Module):
    def __init__(self, mlps=((0, 64, 64, 128), (128, 128, 128, 256), (256, 128, 128, 256)), radii=(0.3, 0.5, 0.7),
                 nsamples=(32, 32, 32), methods=("fps", "sequence", "sequence")):
        super().__init__()
        self.SA_modules = nn.ModuleList(
            m.PointnetSAModuleVotes(mlp=list(mlp), radius=r, nsample=ns, use_xyz=True, normalize_xyz=True, sample_method=meth)
            for mlp, r, ns, meth in zip(mlps, radii, nsamples, methods))
        self.cov_final = nn.Conv1d(mlps[-1][-1], 256, kernel_size=1)

    def forward(self, pts, npoints):
        xyz, feats = pts.contiguous(), None
        for sa, n in zip(self.SA_modules, npoints):
            xyz, feats, _ = sa(xyz, feats, n)
        return xyz, self.cov_final(feats)


class _Head(nn.Module):
    def __init__(self, with_sa):
        super().__init__()
        self.transformer_block = m.TransformerBlock(256, 512, 16)
        if with_sa:
            self.vote_aggregation = m.PointnetSAModuleVotes(mlp=[257, 256, 256, 256], radius=0.3, nsample=16, use_xyz=True,
                                                            normalize_xyz=True, sample_method="fps")


class HotPathNet(nn.Module):
    """forward(search (B,Ns,3), template (B,Nt,3)) -> the tensors HotPath.forward returns (same glue between stages)."""

    def __init__(self, npoints_search=(512, 256, 128), npoints_template=(256, 128, 64), box_npoint=64):
        super().__init__()
        self.backbone_3d = _Backbone()
        self.centroid_voting_head = _Head(with_sa=False)
        self.box_voting_head = _Head(with_sa=True)
        self.npoints_search, self.npoints_template, self.box_npoint = npoints_search, npoints_template, box_npoint

    def forward(self, search, template):
        if self.training:
            from . import train_ops
            train_ops.repack_stale(self.modules())       # every layer's packed images after the optimiser update: one launch
        s_xyz, s_feat = self.backbone_3d(search, self.npoints_search)
        t_xyz, t_feat = self.backbone_3d(template, self.npoints_template)
        cen, _ = self.centroid_voting_head.transformer_block(s_xyz, s_feat.transpose(1, 2).contiguous())
        votes_feats = torch.cat([torch.full_like(cen[:, :, :1], 0.5), cen], dim=2).transpose(1, 2).contiguous()
        b_xyz, b_feat, _ = self.box_voting_head.vote_aggregation(s_xyz, votes_feats, self.box_npoint)
        box, _ = self.box_voting_head.transformer_block(b_xyz, b_feat.transpose(1, 2).contiguous())
        return {"search_feats": s_feat, "template_feats": t_feat, "centroid_feats": cen, "box_sa_feats": b_feat, "box_feats": box}


def time_train_step(device, world=1, rank=0, local_rank=0, batch=48, steps=5, warmup=3, n_search=1024, n_template=512,
                    graph=True):
    """One training step of the hot path the way the reference trains (tools/train_utils/train_utils.py:40-51,
    tools/train_tracking.py:158-159, ptt.yaml OPTIMIZATION): forward in train() mode (BatchNorm batch statistics), loss,
    backward, DistributedDataParallel gradient all-reduce over NCCL (the only collective; needs an initialised process
    group when world > 1), clip_grad_norm_(10), Adam(lr 1e-3, betas (0.5, 0.999)).  Every rank trains on its own `batch`
    synthetic frames (weak scaling).  Timed on the device with CUDA events.  Returns a dict (ms_per_step of THIS rank).

    graph=True: after the eager warm-up the WHOLE step (forward, backward, the NCCL all-reduce DDP issues from its hooks,
    clipping, Adam) is captured once into a CUDA graph and the timed steps are replays with fresh inputs copied into the
    static input buffers.  The step is ~670 launches whose host side (autograd, wrappers, DDP bookkeeping) costs 14 ms
    against 17 ms of kernels on an idle host and MORE than the kernels once 8 ranks share 16 cores; the replay has no host
    side.  Set-up follows torch's rules for capturing a full backward under DDP (async error handling off before
    init_process_group -- see `prepare_env_for_graphs` --, DDP built on a side stream, >= 11 eager iterations first).
    graph=False, or a failed capture (reported in `launch_mode`), times eager steps."""
    import time

    import torch.distributed as dist

    from . import synth

    net = HotPathNet()
    synth.load_filled(net, seed=0)
    net = net.to(device).train()
    side = torch.cuda.Stream(device)
    side.wait_stream(torch.cuda.current_stream(device))
    # DDP as train_tracking.py:158-159 builds it, with three measured settings (2 GPUs: 18.8 -> 18.2 ms per eager step):
    # 4 MB buckets (the 16.4 MB of gradients are all-reduced in pieces while the backward pass is still running),
    # static_graph (same parameters every step), and no per-step buffer broadcast: the BatchNorm running statistics are
    # per-GPU in the reference too (no --sync_bn), rank 0 -- whose buffers a checkpoint holds -- never receives anything,
    # so the broadcast only overwrites the other ranks' copies and costs 0.5 ms per step.
    with torch.cuda.stream(side):
        model = (nn.parallel.DistributedDataParallel(net, device_ids=[local_rank], bucket_cap_mb=4, gradient_as_bucket_view=True,
                                                     broadcast_buffers=False, static_graph=True)
                 if world > 1 else net)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, betas=(0.5, 0.999), capturable=bool(graph))
    n_params = sum(p.numel() for p in net.parameters())
    sets = [(torch.from_numpy(synth.make_clouds(batch, n_search, 3000 + 16 * rank + i, "dense")).to(device),
             torch.from_numpy(synth.make_clouds(batch, n_template, 4000 + 16 * rank + i, "dense", role="template")).to(device))
            for i in range(4)]
    in_s, in_t = torch.empty_like(sets[0][0]), torch.empty_like(sets[0][1])      # static inputs of the captured step

    def step():
        out = model(in_s, in_t)
        loss = sum((v.float() ** 2).mean() for v in out.values())      # synthetic L2 loss on every block output
        opt.zero_grad(set_to_none=True)
        loss.backward()                                                   # DDP all-reduces the gradients in here
        nn.utils.clip_grad_norm_(model.parameters(), 10.0)
        opt.step()
        return loss

    def feed(i):
        in_s.copy_(sets[i % 4][0])
        in_t.copy_(sets[i % 4][1])

    n_warm = max(warmup, 11 if (graph and world > 1) else 3)
    with torch.cuda.stream(side):
        for i in range(n_warm):
            feed(i)
            step()
    torch.cuda.current_stream(device).wait_stream(side)
    torch.cuda.synchronize(device)
    mode, g, g_loss = "eager launches", None, None
    if graph:
        try:
            g = torch.cuda.CUDAGraph()
            opt.zero_grad(set_to_none=True)
            with torch.cuda.graph(g):
                g_loss = step()
            mode = "one CUDA graph per step (forward + backward + all-reduce + clip + Adam), replayed"
        except Exception as e:                                            # noqa: BLE001 - report and time eager steps
            g = None
            torch.cuda.synchronize(device)
            mode = "eager launches (graph capture failed: %s)" % (str(e).splitlines()[0][:120] if str(e) else type(e).__name__)

    def run(i):
        feed(i)
        if g is not None:
            g.replay()
            return g_loss
        return step()

    for i in range(2):
        run(i)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    losses = []
    e0.record()
    t0 = time.perf_counter()
    for i in range(steps):
        losses.append(run(i).detach().clone())
    host_ms = (time.perf_counter() - t0) * 1e3 / steps                  # host time to ENQUEUE a step (no synchronisation inside)
    e1.record()
    torch.cuda.synchronize(device)
    if g is not None:
        from . import train_ops
        train_ops.invalidate_packs(net.modules())     # replays moved the parameters behind the caches' back (see its docstring)
    in_sync = None
    if world > 1:
        # the DDP invariant after the timed steps: every rank holds bit-identical parameters (the all-reduce -- captured
        # in the graph or not -- delivered the same averaged gradients everywhere)
        digest = torch.stack([p.detach().double().sum() for p in net.parameters()] +
                             [p.detach().double().abs().sum() for p in net.parameters()])
        gathered = [torch.empty_like(digest) for _ in range(world)]
        dist.all_gather(gathered, digest)
        in_sync = all(bool(torch.equal(gathered[0], t)) for t in gathered[1:])
    return {"ms_per_step": e0.elapsed_time(e1) / steps, "host_enqueue_ms_per_step": host_ms, "launch_mode": mode,
            "ranks_hold_identical_parameters": in_sync, "steps": steps,
            "warmup": n_warm, "batch_per_gpu": batch, "parameters": n_params,
            "allreduce_bytes_per_step": 4 * n_params if world > 1 else 0,
            "loss_first_last": [float(losses[0]), float(losses[-1])]}


def prepare_env_for_graphs():
    """Call before torch.distributed.init_process_group: capturing NCCL collectives into a CUDA graph needs the process
    group's asynchronous error handling (a watchdog that queries events of in-flight work) switched off."""
    import os

    os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")
    os.environ.setdefault("NCCL_ASYNC_ERROR_HANDLING", "0")
