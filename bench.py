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
    if n == 1024 and a.ntemplate == 512:
        return None
    # NPOINTS scale with N (reproduces the yaml at N = 1024)
    nt = a.ntemplate
    return dict(npoints_search=(n // 2, n // 4, n // 8), npoints_template=(nt // 2, nt // 4, nt // 8))


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
    flops["box.sa.mlp"] = 2.0 * 64 * 16 * mlp_macs([260, 256, 256, 256])

    def tr(n, k=16, dp=256, dm=512):
        return 2.0 * (n * dp * dm + 3 * n * dm * dm + n * k * (3 * dm + dm * dm) + 2 * n * k * dm * dm + n * dm * dp)

    flops["centroid.transformer"] = tr(cfgs["npoints_search"][2])
    flops["box.transformer"] = tr(64)
    return flops


def algorithmic_bytes(a):
    """Per-frame algorithmic bytes of the index kernels: FPS B*(12N + 4M + 12M for the emitted centres),
    ball query B*(12N + 12M + 4*M*ns)."""
    cfgs = scaled_cfg(a) or dict(npoints_search=(512, 256, 128), npoints_template=(256, 128, 64))
    out = {}
    for tag, n0, npts in (("search", a.nsearch, cfgs["npoints_search"]), ("template", a.ntemplate, cfgs["npoints_template"])):
        n = n0
        for l in range(3):
            m = npts[l]
            if l == 0:
                out["%s.sa1.fps" % tag] = 12.0 * n + 16.0 * m
            out["%s.sa%d.ball_query" % (tag, l + 1)] = 12.0 * n + 12.0 * m + 4.0 * m * 32
            n = m
    ns3 = cfgs["npoints_search"][2]
    out["box.sa.fps"] = 12.0 * ns3 + 16.0 * 64
    out["box.sa.ball_query"] = 12.0 * ns3 + 12.0 * 64 + 4.0 * 64 * 16
    return out


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index, period=0.004):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
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
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s), **({"error": self.error} if self.error else {})}


def physical_gpu_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


def cpu_hot_path(a, frames, threads, steps, warmup):
    """The reference's CPU path for the hot path = oracle.torch_port (PyTorch CPU + C ops), timed on host cores."""
    import torch

    from oracle import torch_port
    from ptt_b200 import synth

    torch.set_num_threads(threads)
    sd = synth.hot_path_state_dict(0)
    search = torch.from_numpy(synth.make_clouds(frames, a.nsearch, 900, a.kind))
    template = torch.from_numpy(synth.make_clouds(frames, a.ntemplate, 901, a.kind, role="template"))
    cfg = scaled_cfg(a)
    with torch.no_grad():
        for _ in range(warmup):
            torch_port.hot_path_frame(sd, search, template, cfg)
        t0 = time.perf_counter()
        for _ in range(steps):
            torch_port.hot_path_frame(sd, search, template, cfg)
        dt = time.perf_counter() - t0
    return frames * steps / dt, dt / steps


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    threads = os.cpu_count() or 1
    frames = a.cpu_frames
    steps, warmup = max(1, min(a.steps, 10)), max(1, min(a.warmup, 2))
    fps, sec = cpu_hot_path(a, frames, threads, steps, warmup)
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": a.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(a, a.gpus),
            "cpu_baseline": {"value": fps, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                             "sample": "%d frames per step x %d steps (the reference is Python over a CUDA-only "
                                       "third-party extension; its CPU path is the oracle port, oracle/torch_port.py "
                                       "over oracle/pointnet2_ref.c)" % (frames, steps)},
            "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_b200(a):
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

    # ---- device-resident throughput: two steps in flight (HostPipeline slots fed from HBM-resident inputs) ----
    # No L2 flush is possible between overlapping steps; instead the rotating input sets together exceed the L2
    # (n_big sets x 0.88 MB > 126 MB), so every step's clouds come from HBM.
    pipe = hotpath.HostPipeline(synth.hot_path_state_dict(0), cfg=scaled_cfg(a), device=dev, depth=2)
    n_big = 160
    big_s = torch.cat([search_d[i % n_sets] for i in range(n_big)]).view(n_big, B, a.nsearch, 3).clone()
    big_t = torch.cat([templ_d[i % n_sets] for i in range(n_big)]).view(n_big, B, a.ntemplate, 3).clone()
    big_s += torch.arange(n_big, device=dev, dtype=torch.float32).view(-1, 1, 1, 1) * 1e-6      # distinct bits per set
    for i in range(max(4, a.warmup)):
        pipe.push(big_s[i % n_big], big_t[i % n_big], to_host=False)
    pipe.drain(to_host=False)
    barrier()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for i in range(a.steps):
        pipe.push(big_s[(7 * i) % n_big], big_t[(7 * i) % n_big], to_host=False, after=p0)
    pipe.drain(to_host=False)
    for slot in pipe.slots:
        torch.cuda.current_stream().wait_stream(slot._io_stream)
    p1.record()
    barrier()
    pipe_ms = p0.elapsed_time(p1)
    clocks = sampler.stop()        # sampled over both device-timed regions (sequential + pipelined)

    # ---- end-to-end through the host API: pinned host in, pinned host out, copies inside the timed region ----
    # HostPipeline = the throughput form of HotPath.forward_host: two instances alternate, so the H2D copy and compute
    # of step i+1 overlap the D2H copy of step i.  Every step's inputs come from pinned host memory and all ten
    # outputs of every step are copied back to pinned host memory inside the timed region.
    for i in range(max(4, a.warmup)):
        pipe.push(search_h[i % n_sets], templ_h[i % n_sets])
    pipe.drain()
    barrier()
    t0 = time.perf_counter()
    got = 0
    for i in range(a.steps):
        got += pipe.push(search_h[i % n_sets], templ_h[i % n_sets]) is not None
    tail = pipe.drain()
    got += len(tail)
    out_h = tail[-1]
    barrier()
    e2e_s = time.perf_counter() - t0
    assert got == a.steps
    h2d = search_h[0].numel() * 4 + templ_h[0].numel() * 4
    d2h = sum(v.numel() * v.element_size() for v in out_h.values())
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

    dev_ms, e2e_ms, wall_ms, pipe_ms = shard.max_over_ranks([dev_ms, e2e_s * 1e3, t_wall * 1e3, pipe_ms], device=dev)   # slowest rank

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
        peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
        peak_gbs = peaks.get("hbm_gbs", 6650.0)
        abytes = algorithmic_bytes(a)
        achieved_tf = alg[top] * B / (known[top] * 1e-3) / 1e12
        traffic = None
        tp = os.path.join(REPO, "profiles", "r1_traffic.json")     # dram__bytes_read+write per launch from the ncu --set full capture
        if os.path.exists(tp) and B == 48 and a.nsearch == 1024:
            traffic = json.load(open(tp)).get(top)
        line = {
            "metric": METRIC, "value": shard.whole_job_throughput(B * a.steps, n_gpus, pipe_ms * 1e-3), "unit": UNIT, "n_gpus": n_gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": pipe_ms / a.steps,
            "value_timing": "K steps, two in flight (depth-2 pipeline of CUDA-graph replays), one CUDA-event pair around all K, "
                            "inputs rotate over %d HBM-resident sets (> L2)" % n_big,
            "sequential_l2_flushed": {"value": shard.whole_job_throughput(B * a.steps, n_gpus, dev_ms * 1e-3), "unit": UNIT,
                                      "ms_per_step": dev_ms / a.steps,
                                      "how": "one step at a time, CUDA events per step, 256 MiB L2 flush between steps"}, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(a, n_gpus),
            "e2e": {"value": frames / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / a.steps, "api": "HostPipeline(depth=2).push/drain",
                    "sync_one_step_at_a_time": {"value": B * a.steps / e2e_sync_s, "unit": UNIT,
                                                "ms_per_step": e2e_sync_s * 1e3 / a.steps, "api": "HotPath.forward_host"}},
            "gpu_launches": launches, "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": top, "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved_tf / peak_tf, "traffic": traffic,
                         "frac_ceiling": "1/3: fp32-class accuracy costs 3 fp16 MMAs per algorithmic MAC (DESIGN.md 4)",
                         "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback",
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
        if whole is not None:
            line["whole_model"] = {"value": shard.whole_job_throughput(B * a.steps, n_gpus, whole[0] * 1e-3), "unit": UNIT,
                                   "ms_per_step": whole[0] / a.steps, "extra_stage_ms": whole[1],
                                   "what": "HotPath.forward_full: the whole tracker forward in eval mode (hot path + CosineSimAug "
                                           "+ cla / vote / refine stacks -> pred_box_data), one CUDA-graph replay at a time, L2 "
                                           "flushed between steps"}
        if not a.no_cpu_baseline and world == 1:
            cpu_threads = os.cpu_count() or 1
            fps, sec = cpu_hot_path(a, a.cpu_frames, cpu_threads, steps=3, warmup=1)
            line["cpu_baseline"] = {"value": fps, "unit": UNIT, "cores": cpu_threads, "kind": "port",
                                    "sample": "%d frames per step x 3 steps of the same workload, oracle/torch_port.py "
                                              "over oracle/pointnet2_ref.c" % a.cpu_frames}
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
