#!/usr/bin/env python
"""bench.py -- frames/s of the PTT point-feature hot path (SA stack + transformer blocks) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is one pass of the hot path (ptt_b200.hotpath.HotPath: backbone SA1-3 on the search and template
clouds + cov_final, centroid-head transformer block, box-head SA, box-head transformer block) over one batch of
synthetic frames.  Workload = BASELINE.json configs[1]: KITTI Car ptt.yaml, N=1024 search / 512 template points,
batch 48 per GPU (weak scaling: every rank processes its own 48 frames; no data-path collective).

One JSON line on stdout (rank 0).  `value`: inputs resident in HBM, device-timed.  `e2e`: through the host API
(HotPath.forward_host) with pinned host buffers, H2D + D2H inside the timed region.  `roofline`: the dominant
stage, algorithmic FLOPs / CUDA-event time on its stream.  `cpu_baseline`: the CPU oracle port on a bounded sample.
"""
import argparse
import json
import os
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

METRIC = "tracked frames/sec (SA + transformer hot path, forward)"
UNIT = "frames/s"
# steps in flight in the pipelined measurements (HostPipeline slots, each a HotPath with its own graph, workspaces and
# streams).  Measured: 2 -> 3 slots +3 % device-resident and +8 % end to end (the third slot's H2D / D2H copies hide
# completely under the other two's kernels); 4 adds < 1 %.
PIPE_DEPTH = 3


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=48, help="frames per GPU per step")
    ap.add_argument("--nsearch", type=int, default=1024)
    ap.add_argument("--ntemplate", type=int, default=512)
    ap.add_argument("--kind", default="dense", choices=["dense", "sparse"])
    ap.add_argument("--cpu-frames", type=int, default=4, help="frames per CPU-baseline step (bounded sample)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-whole-model", action="store_true", help="skip the whole-tracker-forward figure")
    ap.add_argument("--no-train-step", action="store_true", help="skip the training-step figure (BASELINE configs[3])")
    ap.add_argument("--no-tracking", action="store_true", help="skip the batched tracking-loop figure (SURVEY 8(f) N3)")
    ap.add_argument("--no-reference-gpu", action="store_true", help="skip the reference-modules-on-this-GPU figure")
    ap.add_argument("--sustain-s", type=float, default=2.5, help="length of the sustained region in seconds (0 = skip)")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of replaying the CUDA graph")
    return ap.parse_args()


def workload_config(a, n_gpus):
    return {"workload": "KITTI Car ptt.yaml hot path: N=%d search + %d template pts, batch %d per GPU, %s synthetic crops"
                        % (a.nsearch, a.ntemplate, a.batch, a.kind),
            "global_batch": a.batch * n_gpus, "batch_per_gpu": a.batch, "n_search": a.nsearch, "n_template": a.ntemplate,
            "mode": "eval (BatchNorm folded), forward", "parallelism": "batch-sharded replicas x%d, no collective" % n_gpus,
            "l2": "inputs larger than L2: 160 rotating HBM-resident input sets (141 MB); the sequential figure flushes L2 "
                  "between steps with a 256 MiB write"}


def scaled_cfg(a):
    n = a.nsearch
    if a.ntemplate == 512 and n in (512, 1024):     # configs 2 and 3: the yaml's NPOINTS unchanged (config 3: SA1 FPS 512 -> 512)
        return None
    # NPOINTS scale with N (reproduces the yaml at N = 1024)
    nt = a.ntemplate
    return dict(npoints_search=(n // 2, n // 4, n // 8), npoints_template=(nt // 2, nt // 4, nt // 8),
                box_npoint=min(64, n // 16))           # the box head samples half of the seeds (64 of 128 in the yaml)


# per-frame algorithmic work of the hot path (DESIGN.md "Measurement")
def algorithmic(a):
    cfgs = scaled_cfg(a) or dict(npoints_search=(512, 256, 128), npoints_template=(256, 128, 64))
    specs = ([3, 64, 64, 128], [131, 128, 128, 256], [259, 128, 128, 256])

    def mlp_macs(spec):
        return sum(spec[i] * spec[i + 1] for i in range(len(spec) - 1))

    flops = {}
    for tag, npts in (("search", cfgs["npoints_search"]), ("template", cfgs["npoints_template"])):
        for l in range(3):
            flops["%s.sa%d.mlp" % (tag, l + 1)] = 2.0 * npts[l] * 32 * mlp_macs(specs[l])
    nbox = cfgs.get("box_npoint", 64)
    flops["box.sa.mlp"] = 2.0 * nbox * 16 * mlp_macs([260, 256, 256, 256])

    def tr(n, k=16, dp=256, dm=512):
        return 2.0 * (n * dp * dm + 3 * n * dm * dm + n * k * (3 * dm + dm * dm) + 2 * n * k * dm * dm + n * dm * dp)

    flops["centroid.transformer"] = tr(cfgs["npoints_search"][2])
    flops["box.transformer"] = tr(nbox)
    return flops


def algorithmic_bytes(a):
    """Per-frame algorithmic bytes of the index kernels: FPS B*(12N + 4M + 12M for the emitted centres),
    ball query B*(12N + 12M + 4*M*ns)."""
    cfgs = scaled_cfg(a) or dict(npoints_search=(512, 256, 128), npoints_template=(256, 128, 64))
    out = {}
    for tag, n0, npts in (("search", a.nsearch, cfgs["npoints_search"]), ("template", a.ntemplate, cfgs["npoints_template"])):
        n = n0
        bq = 0.0
        for l in range(3):
            m = npts[l]
            if l == 0:
                out["%s.sa1.fps" % tag] = 12.0 * n + 16.0 * m
            bq += 12.0 * n + 12.0 * m + 4.0 * m * 32
            n = m
        out["%s.ball_query" % tag] = bq            # the three levels of a branch are ONE launch (ptt_ball_query_nested)
    ns3 = cfgs["npoints_search"][2]
    nbox = cfgs.get("box_npoint", 64)
    out["box.sa.fps"] = 12.0 * ns3 + 16.0 * nbox
    out["box.sa.ball_query"] = 12.0 * ns3 + 12.0 * nbox + 4.0 * nbox * 16
    return out


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index, period=0.004):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz, self.power = [], set(), None, []
        self._halt = threading.Event()
        self.error = None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                     "hw_power_brake_slowdown": 0x80, "sync_boost": 0x10, "applications_clocks_setting": 0x2}
            while not self._halt.is_set():
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    self.power.append(nv.nvmlDeviceGetPowerUsage(h) / 1000.0)
                except Exception:
                    pass
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for n, bit in names.items():
                    if mask & bit:
                        self.reasons.add(n)
                time.sleep(self.period)
        except Exception as e:  # NVML missing: report it, do not fail the bench
            self.error = repr(e)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        pw = sorted(self.power)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s), "power_w_median": pw[len(pw) // 2] if pw else None, "power_w_max": pw[-1] if pw else None,
                **({"error": self.error} if self.error else {})}


def physical_gpu_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


def cpu_hot_path(a, frames, threads, steps, warmup):
    """The reference's CPU implementation of the hot path, timed on the host cores: the reference's OWN nn.Modules
    (oracle/_ref staged by oracle/make_ref.sh, or /root/reference) driven by oracle/ref_hotpath.py when the tree is
    there -- its CUDA-only third-party pointnet2_ops replaced by the oracle's C ops -- else the oracle port.
    Returns (frames/s, s/step, kind, description)."""
    import torch

    from oracle import refload, torch_port
    from ptt_b200 import synth

    torch.set_num_threads(threads)
    search = torch.from_numpy(synth.make_clouds(frames, a.nsearch, 900, a.kind))
    template = torch.from_numpy(synth.make_clouds(frames, a.ntemplate, 901, a.kind, role="template"))
    cfg = scaled_cfg(a)
    if refload.available() and cfg is None:
        from oracle import ref_hotpath
        ref = ref_hotpath.RefHotPath(synth.full_model_state_dict(0), "cpu")
        fn = lambda: ref.hot_path(search, template)
        kind, what = "reference", ("the reference's own ptt.models modules (%s) on the host cores, eval mode; its CUDA-only "
                                   "pointnet2_ops replaced by oracle/pointnet2_ref.c" % os.path.relpath(refload.REFERENCE_ROOT, REPO))
    else:
        sd = synth.hot_path_state_dict(0)
        fn = lambda: torch_port.hot_path_frame(sd, search, template, cfg)
        kind, what = "port", "oracle/torch_port.py over oracle/pointnet2_ref.c"
    with torch.no_grad():
        for _ in range(warmup):
            fn()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        dt = time.perf_counter() - t0
    return frames * steps / dt, dt / steps, kind, what


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    threads = os.cpu_count() or 1
    frames = a.cpu_frames
    steps, warmup = max(1, min(a.steps, 10)), max(1, min(a.warmup, 2))
    fps, sec, kind, what = cpu_hot_path(a, frames, threads, steps, warmup)
    cfg = workload_config(a, a.gpus)
    # what this arm actually ran: a bounded sample of the workload (the CPU needs ~0.03 s per frame)
    cfg["reference_sample"] = {"frames_per_step": frames, "steps": steps, "warmup": warmup, "host_threads": threads,
                               "note": "batch_per_gpu above is the GPU arm's; this arm times %d frames per step" % frames}
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": a.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": fps, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
                             "sample": "%d frames per step x %d steps: %s" % (frames, steps, what)},
            "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_b200(a):
    import numpy as np
    import torch
    import torch.distributed as dist

    from ptt_b200 import _lib, hotpath, shard, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        from ptt_b200 import train as _train_env
        _train_env.prepare_env_for_graphs()          # the training step captures its NCCL all-reduce into a CUDA graph
        dist.init_process_group("nccl", device_id=dev)
    n_gpus = world

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    hp = hotpath.HotPath(synth.hot_path_state_dict(0), cfg=scaled_cfg(a), device=dev)
    B = a.batch
    # each rank owns its own shard of the global batch (distinct seeds): rank r gets frames [r*B, (r+1)*B)
    n_sets = 4
    search_h = [torch.from_numpy(synth.make_clouds(B, a.nsearch, 1000 + 16 * rank + i, a.kind)).pin_memory() for i in range(n_sets)]
    templ_h = [torch.from_numpy(synth.make_clouds(B, a.ntemplate, 2000 + 16 * rank + i, a.kind, role="template")).pin_memory()
               for i in range(n_sets)]
    search_d = [x.to(dev) for x in search_h]
    templ_d = [x.to(dev) for x in templ_h]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step(i):
        if a.no_graph:
            return hp(search_d[i % n_sets], templ_d[i % n_sets])
        return hp.forward_graph(search_d[i % n_sets], templ_d[i % n_sets])

    for i in range(a.warmup):
        step(i)
    barrier()

    # ---- device-resident timed region: K steps, each bracketed by events, L2 flushed between steps ----
    sampler = ClockSampler(physical_gpu_index(local_rank))
    sampler.start()
    launches0 = _lib.launch_count()
    ev = []
    barrier()
    t_wall0 = time.perf_counter()
    for i in range(a.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step(i)
        e1.record()
        ev.append((e0, e1))
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = _lib.launch_count() - launches0
    dev_ms = sum(x.elapsed_time(y) for x, y in ev)
    if not a.no_graph:
        # graph replays do not pass through the library's host entry points: count what one replay launches
        l0 = _lib.launch_count()
        hp(search_d[0], templ_d[0])
        launches = (_lib.launch_count() - l0) * a.steps

    # ---- per-stage pass (roofline): the same K steps launched eagerly with a CUDA-event pair around every stage ----
    hp.profile(True)
    hp.overlap = False            # one stream: a stage's events then bracket only its own kernels
    barrier()
    for i in range(a.steps):
        flush.zero_()
        hp(search_d[i % n_sets], templ_d[i % n_sets])
    barrier()
    stage_ms = hp.stage_ms(median=True)
    hp.profile(False)
    hp.overlap = True

    # ---- device-resident throughput: PIPE_DEPTH steps in flight (HostPipeline slots fed from HBM-resident inputs) ----
    # No L2 flush is possible between overlapping steps; instead the rotating input sets together exceed the L2
    # (n_big sets x 0.88 MB > 126 MB), so every step's clouds come from HBM.
    pipe = hotpath.HostPipeline(synth.hot_path_state_dict(0), cfg=scaled_cfg(a), device=dev, depth=PIPE_DEPTH)
    n_big = 160
    big_s = torch.cat([search_d[i % n_sets] for i in range(n_big)]).view(n_big, B, a.nsearch, 3).clone()
    big_t = torch.cat([templ_d[i % n_sets] for i in range(n_big)]).view(n_big, B, a.ntemplate, 3).clone()
    big_s += torch.arange(n_big, device=dev, dtype=torch.float32).view(-1, 1, 1, 1) * 1e-6      # distinct bits per set
    for i in range(max(4, a.warmup)):
        pipe.push(big_s[i % n_big], big_t[i % n_big], to_host=False)
    pipe.drain()
    barrier()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for i in range(a.steps):
        pipe.push(big_s[(7 * i) % n_big], big_t[(7 * i) % n_big], to_host=False, after=p0)
    pipe.drain()
    for slot in pipe.slots:
        torch.cuda.current_stream().wait_stream(slot._io_stream)
    p1.record()
    barrier()
    pipe_ms = p0.elapsed_time(p1)
    clocks = sampler.stop()        # sampled over both device-timed regions (sequential + pipelined)

    # ---- sustained: the same pipelined loop for >= a.sustain_s seconds of back-to-back steps (the K-step region above is
    # a ~30 ms burst: clocks and power have not settled there), clocks / power / throttle reasons sampled throughout
    sustained = None
    if a.sustain_s > 0:
        n_sus = max(a.steps, int(a.sustain_s * 1e3 / (pipe_ms / a.steps)) + 1)
        sus_sampler = ClockSampler(physical_gpu_index(local_rank), period=0.02)
        barrier()
        sus_sampler.start()
        q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        q0.record()
        for i in range(n_sus):
            pipe.push(big_s[(7 * i) % n_big], big_t[(7 * i) % n_big], to_host=False, after=q0)
        pipe.drain()
        for slot in pipe.slots:
            torch.cuda.current_stream().wait_stream(slot._io_stream)
        q1.record()
        barrier()
        sus_ms = q0.elapsed_time(q1)
        sustained = (n_sus, shard.max_over_ranks([sus_ms], device=dev)[0], sus_sampler.stop())

    # ---- end-to-end through the host API: pinned host in, pinned host out, copies inside the timed region ----
    # HostPipeline = the throughput form of HotPath.forward_host: two instances alternate, so the H2D copy and compute
    # of step i+1 overlap the D2H copy of step i.  Every step's inputs come from pinned host memory and the selected
    # outputs of every step are copied back to pinned host memory inside the timed region.  The default selection of
    # the hot path is its final stage's features (box_feats: what the refine stack consumes); `e2e.all_outputs` repeats
    # the measurement with every output tensor of the step copied back (round 1's figure).
    def e2e_run(p):
        for i in range(max(4, a.warmup)):
            p.push(search_h[i % n_sets], templ_h[i % n_sets])
        p.drain()
        barrier()
        t0 = time.perf_counter()
        got = 0
        for i in range(a.steps):
            got += p.push(search_h[i % n_sets], templ_h[i % n_sets]) is not None
        tail = p.drain()
        got += len(tail)
        barrier()
        dt = time.perf_counter() - t0
        assert got == a.steps
        return dt, sum(v.numel() * v.element_size() for v in tail[-1].values()), sorted(tail[-1])

    e2e_s, d2h, e2e_keys = e2e_run(pipe)
    pipe.outputs = "all"
    e2e_all_s, d2h_all, _ = e2e_run(pipe)
    pipe.outputs = None
    h2d = search_h[0].numel() * 4 + templ_h[0].numel() * 4
    # latency form (one step at a time, synchronous): reported next to the throughput form
    for i in range(2):
        hp.forward_host(search_h[i % n_sets], templ_h[i % n_sets])
    barrier()
    t0 = time.perf_counter()
    for i in range(a.steps):
        hp.forward_host(search_h[i % n_sets], templ_h[i % n_sets])
    barrier()
    e2e_sync_s = time.perf_counter() - t0

    # ---- whole tracker forward (hot path + CosineSimAug + the heads' Conv1d stacks; SURVEY.md 8(d) "also report"):
    # sequential CUDA-graph replays, events per step, L2 flushed between steps
    whole = None
    if not a.no_whole_model:
        hpf = hotpath.HotPath(synth.full_model_state_dict(0), cfg=scaled_cfg(a), device=dev)
        for i in range(max(3, a.warmup)):
            hpf.forward_graph(search_d[i % n_sets], templ_d[i % n_sets], full=True)
        barrier()
        evw = []
        for i in range(a.steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            hpf.forward_graph(search_d[i % n_sets], templ_d[i % n_sets], full=True)
            e1.record()
            evw.append((e0, e1))
        barrier()
        whole_ms = sum(x.elapsed_time(y) for x, y in evw)
        hpf.profile(True)
        hpf.overlap = False
        for i in range(min(a.steps, 5)):
            hpf.forward_full(search_d[i % n_sets], templ_d[i % n_sets])
        barrier()
        wst = hpf.stage_ms(median=True)
        whole = (shard.max_over_ranks([whole_ms], device=dev)[0], {k: round(v, 4) for k, v in sorted(wst.items())
                                                                    if k.startswith(("cosine", "centroid.heads", "box.heads"))})
        del hpf

    # ---- the reference's OWN modules on this GPU over the `pointnet2_ops._ext` drop-in (Level-1 integration): torch /
    # cuDNN / cuBLAS arithmetic exactly as the reference would run it on a B200 with a working pointnet2_ops -- the
    # "unfused" comparison point for the fused path (rank 0 of a single-GPU run only; needs oracle/_ref)
    ref_gpu = None
    if world == 1 and not a.no_reference_gpu and scaled_cfg(a) is None:
        from oracle import refload
        if refload.available():
            from oracle import ref_hotpath
            ref = ref_hotpath.RefHotPath(synth.full_model_state_dict(0), dev)
            for i in range(3):
                ref.hot_path(search_d[i % n_sets], templ_d[i % n_sets])
            torch.cuda.synchronize()
            r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n_ref = max(3, min(a.steps, 10))
            r0.record()
            for i in range(n_ref):
                ref.hot_path(search_d[i % n_sets], templ_d[i % n_sets])
            r1.record()
            torch.cuda.synchronize()
            ref_gpu = {"value": B * n_ref / (r0.elapsed_time(r1) * 1e-3), "unit": UNIT, "steps": n_ref,
                       "ms_per_step": r0.elapsed_time(r1) / n_ref,
                       "what": "the reference's own ptt.models modules (oracle/_ref) on this B200, eval mode, torch defaults "
                               "(cuDNN / cuBLAS TF32 allowed), their pointnet2_ops._ext calls served by libptt_b200.so; "
                               "inputs resident in HBM, eager launches as the reference issues them"}
            del ref

    # ---- the tracking loop itself.  T independent synthetic tracklets advance in lockstep; per frame the
    # raw clouds (pinned host) are copied H2D, cropped around the previous boxes, resampled, run through the whole tracker
    # forward and the boxes updated -- all inside one CUDA graph (ptt_b200.tracking.BatchedTracker).  "tracked frames/s" =
    # T * frames / wall time incl. the H2D copies.  The reference's loop is T = 1 with numpy pre / post-processing on the
    # host and a host round trip per frame: `host_prepost_b1` times that structure with OUR model at batch 1 (the oracle's
    # numpy restatement of the reference's crop / regularize / box update as the host part -- a baseline leg).
    track = None
    if not a.no_tracking and scaled_cfg(a) is None:
        from ptt_b200 import synth_tracks, tracking
        track = {}
        n_frames, cap = 17, 5200
        sd_full = synth.full_model_state_dict(0)
        for T in (1, 8, 64):
            tracks = synth_tracks.make_tracklets(T, n_frames, seed=40 + rank, points_per_frame=4200)
            frames = [tuple(torch.from_numpy(x).pin_memory() for x in synth_tracks.pad_frames(tracks, i, cap)) for i in range(n_frames)]
            first = torch.from_numpy(np.stack([b[0].as_row() for _, b in tracks]))
            trk = tracking.BatchedTracker(sd_full, T, cap, n_frames, device=dev)
            trk.reset(frames[0][0], frames[0][1], first)
            tracking.run_tracklets(trk, frames[1:5])              # warm-up incl. graph capture
            trk.reset(frames[0][0], frames[0][1], first)
            barrier()
            t0 = time.perf_counter()
            res = tracking.run_tracklets(trk, frames[1:])
            dt = time.perf_counter() - t0
            assert bool(torch.isfinite(res).all())
            track["T=%d" % T] = {"tracked_frames_per_s": T * (n_frames - 1) / dt, "ms_per_frame_step": dt * 1e3 / (n_frames - 1),
                                 "h2d_bytes_per_frame_step": frames[1][0].numel() * 4 + frames[1][1].numel() * 4}
            if T == 1 and world == 1 and not a.no_cpu_baseline:
                from oracle import tracking_ref
                hp1 = hotpath.HotPath(sd_full, device=dev)
                clouds, boxes = tracks[0]

                def model(sr, tm):
                    o = hp1.forward_host(torch.from_numpy(sr).pin_memory(), torch.from_numpy(tm).pin_memory(),
                                         outputs=("pred_box_data",), full=True)
                    return o["pred_box_data"][0].numpy()
                tracking_ref.track(clouds[:4], boxes[0], model)
                t0 = time.perf_counter()
                tracking_ref.track(clouds, boxes[0], model)
                dt1 = time.perf_counter() - t0
                track["host_prepost_b1"] = {"tracked_frames_per_s": (n_frames - 1) / dt1, "ms_per_frame": dt1 * 1e3 / (n_frames - 1),
                                            "what": "the reference's loop structure: numpy crop / regularize / box update on the "
                                                    "host per frame (oracle/tracking_ref.py), HotPath.forward_host(full=True) at "
                                                    "batch 1 with a host round trip per frame"}
                del hp1
            del trk
        track["per_gpu"] = True          # every rank advances its own T tracklets; the figures are rank 0's
        track["what"] = ("BatchedTracker: %d frames per tracklet, ~4200 raw points per frame (padded to %d), crop + resample + "
                         "whole tracker forward + box update in one CUDA graph per frame, H2D of the raw clouds staged one frame "
                         "ahead; wall clock" % (n_frames - 1, cap))

    # ---- BASELINE configs[3] / SURVEY.md 8(e): the training step with the DDP gradient all-reduce -- the one collective the
    # path has.  Every rank runs it (so the driver's 1/2/4/8-GPU scaling run records the NCCL all-reduce), max over ranks.
    train = None
    if not a.no_train_step and scaled_cfg(a) is None:
        from ptt_b200 import train as train_mod
        pipe = None
        torch.cuda.empty_cache()
        train = train_mod.time_train_step(dev, world, rank, local_rank, batch=B, steps=max(3, min(a.steps, 5)), warmup=3,
                                          n_search=a.nsearch, n_template=a.ntemplate)
        keys = [k for k in ("ms_per_step", "host_enqueue_ms_per_step")
                if k in train]
        for k, v in zip(keys, shard.max_over_ranks([train[k] for k in keys], device=dev)):   # slowest rank
            train[k] = v

    dev_ms, e2e_ms, wall_ms, pipe_ms, e2e_all_ms = shard.max_over_ranks([dev_ms, e2e_s * 1e3, t_wall * 1e3, pipe_ms, e2e_all_s * 1e3],
                                                                        device=dev)   # slowest rank

    if rank == 0:
        frames = B * n_gpus * a.steps
        alg = algorithmic(a)
        # dominant stage = the one with the most time per step
        known = {k: v for k, v in stage_ms.items() if k in alg}
        top = max(known, key=known.get)
        peaks = {}
        pk = os.path.join(REPO, "MEASURED_PEAKS.json")
        if os.path.exists(pk):
            peaks = json.load(open(pk))
        # the per-stage times come from a short eager pass (tens of ms: burst clocks) -> the burst peak is the matching
        # denominator; the sustained region is compared with the sustained peak
        peak_tf = peaks.get("bf16_tflops", 1650.0)
        peak_tf_sus = peaks.get("bf16_tflops_sustained", 1400.0)
        peak_gbs = peaks.get("hbm_gbs", 6650.0)
        abytes = algorithmic_bytes(a)
        achieved_tf = alg[top] * B / (known[top] * 1e-3) / 1e12
        traffic = None
        # dram__bytes_read + write of the stage's kernels from the committed `ncu --set full` capture of this workload
        # (profiles/summarize.py writes the json; it cannot be measured inside an un-profiled run)
        traffic_src = None
        for name in ("r2_traffic.json", "r1_traffic.json"):
            tp = os.path.join(REPO, "profiles", name)
            if os.path.exists(tp) and B == 48 and a.nsearch == 1024 and a.kind == "dense":
                traffic = json.load(open(tp)).get(top)
                traffic_src = "profiles/" + name
                break
        line = {
            "metric": METRIC, "value": shard.whole_job_throughput(B * a.steps, n_gpus, pipe_ms * 1e-3), "unit": UNIT, "n_gpus": n_gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": pipe_ms / a.steps,
            "value_timing": "K steps, %d in flight (depth-%d pipeline of CUDA-graph replays), one CUDA-event pair around all K, " % (PIPE_DEPTH, PIPE_DEPTH) +
                            "inputs rotate over %d HBM-resident sets (> L2)" % n_big,
            "sequential_l2_flushed": {"value": shard.whole_job_throughput(B * a.steps, n_gpus, dev_ms * 1e-3), "unit": UNIT,
                                      "ms_per_step": dev_ms / a.steps,
                                      "how": "one step at a time, CUDA events per step, 256 MiB L2 flush between steps"}, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(a, n_gpus),
            "e2e": {"value": frames / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / a.steps, "api": "HostPipeline(depth=%d).push/drain" % PIPE_DEPTH, "outputs": e2e_keys,
                    "all_outputs": {"value": frames / (e2e_all_ms * 1e-3), "unit": UNIT, "d2h_bytes_per_step": d2h_all,
                                    "ms_per_step": e2e_all_ms / a.steps},
                    "sync_one_step_at_a_time": {"value": B * a.steps / e2e_sync_s, "unit": UNIT,
                                                "ms_per_step": e2e_sync_s * 1e3 / a.steps, "api": "HotPath.forward_host"}},
            "gpu_launches": launches, "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": top, "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved_tf / peak_tf, "traffic": traffic, "traffic_source": traffic_src,
                         "frac_of_sustained_peak": achieved_tf / peak_tf_sus, "peak_sustained": peak_tf_sus,
                         "frac_ceiling": "1/3: fp32-class accuracy costs 3 fp16 MMAs per algorithmic MAC (DESIGN.md 4)",
                         "peak_source": ("MEASURED_PEAKS.json bf16_tflops (burst: the stage is timed in a short eager pass)"
                                         if peaks else "fallback"),
                         "flops_per_launch": alg[top] * B, "ms_per_launch": known[top]},
            "stage_ms": {k: round(v, 4) for k, v in sorted(stage_ms.items())},
            # every stage against ITS roofline (same eager pass): dense stages vs the measured bf16 peak (ceiling 1/3, see
            # frac_ceiling), index kernels vs the measured HBM copy bandwidth (they are latency-bound: serial arg-max chain /
            # ordered scan over <= 12 KB per cloud, so the HBM fraction is tiny by construction)
            "stage_roofline": {**{k: {"bound": "tensor", "achieved_tflops": round(alg[k] * B / (v * 1e-3) / 1e12, 1),
                                      "frac": round(alg[k] * B / (v * 1e-3) / 1e12 / peak_tf, 4)}
                                  for k, v in sorted(stage_ms.items()) if k in alg},
                               **{k: {"bound": "hbm", "achieved_gbs": round(abytes[k] * B / (v * 1e-3) / 1e9, 2),
                                      "frac": round(abytes[k] * B / (v * 1e-3) / 1e9 / peak_gbs, 5)}
                                  for k, v in sorted(stage_ms.items()) if k in abytes}},
            "launch_mode": "eager" if a.no_graph else "CUDA graph replay (stage_ms / roofline from an eager single-stream pass of the same steps)",
            "wall_ms_per_step_incl_flush": wall_ms / a.steps,
        }
        if sustained is not None:
            n_sus, sus_ms, sus_clk = sustained
            step_flops = sum(alg.values()) * B
            tf = step_flops / (sus_ms / n_sus * 1e-3) / 1e12
            line["sustained"] = {"value": shard.whole_job_throughput(B * n_sus, n_gpus, sus_ms * 1e-3), "unit": UNIT, "steps": n_sus,
                                 "seconds": sus_ms * 1e-3, "ms_per_step": sus_ms / n_sus, "clocks": sus_clk,
                                 "whole_step_tflops": tf, "whole_step_frac_of_sustained_peak": tf / peak_tf_sus,
                                 "whole_step_frac_of_burst_peak": tf / peak_tf,
                                 "how": "same depth-%d pipeline of graph replays as `value`, back to back for >= %.1f s" % (PIPE_DEPTH, a.sustain_s)}
        if ref_gpu is not None:
            line["reference_modules_gpu"] = ref_gpu
        if track is not None:
            line["tracking"] = track
        if train is not None:
            line["train_step"] = {"value": shard.whole_job_throughput(B, n_gpus, train["ms_per_step"] * 1e-3), "unit": UNIT,
                                  **train, "n_gpus": n_gpus, "scaling": "weak",
                                  "collective": ("NCCL gradient all-reduce (DistributedDataParallel), %d bytes per step"
                                                 % train["allreduce_bytes_per_step"]) if n_gpus > 1 else "none (1 GPU)",
                                  "what": "forward in train() mode (BatchNorm batch statistics) + L2 loss + backward + DDP gradient "
                                          "all-reduce + clip_grad_norm_(10) + Adam, as train_utils.py:40-51 / ptt.yaml OPTIMIZATION; "
                                          "CUDA events, max over ranks"}
        if whole is not None:
            line["whole_model"] = {"value": shard.whole_job_throughput(B * a.steps, n_gpus, whole[0] * 1e-3), "unit": UNIT,
                                   "ms_per_step": whole[0] / a.steps, "extra_stage_ms": whole[1],
                                   "what": "HotPath.forward_full: the whole tracker forward in eval mode (hot path + CosineSimAug "
                                           "+ cla / vote / refine stacks -> pred_box_data), one CUDA-graph replay at a time, L2 "
                                           "flushed between steps"}
        if not a.no_cpu_baseline and world == 1:
            cpu_threads = os.cpu_count() or 1
            fps, sec, kind, what = cpu_hot_path(a, a.cpu_frames, cpu_threads, steps=3, warmup=1)
            line["cpu_baseline"] = {"value": fps, "unit": UNIT, "cores": cpu_threads, "kind": kind,
                                    "sample": "%d frames per step x 3 steps of the same workload: %s" % (a.cpu_frames, what)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
