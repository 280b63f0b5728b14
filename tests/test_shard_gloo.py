"""CPU, world_size 2 over gloo: the host-side sharding / timing-reduction logic bench.py uses for N > 1."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ptt_b200 import shard


def test_shard_range_partitions_the_batch():
    for total in (0, 1, 7, 48, 100):
        for world in (1, 2, 3, 8):
            spans = [shard.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        batch = {"search": torch.arange(10 * 4 * 3, dtype=torch.float32).reshape(10, 4, 3), "ids": torch.arange(10)}
        mine = shard.shard_batch(batch, rank, world)
        # every frame is owned by exactly one rank: gather the ids and check the partition
        sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([mine["ids"].numel()]))
        gathered = [torch.zeros(int(s), dtype=torch.int64) for s in sizes]
        dist.all_gather(gathered, mine["ids"]) if len({int(s) for s in sizes}) == 1 else None
        # max-over-ranks timing: rank r reports (r + 1) ms -> everyone sees world ms
        t_max, = shard.max_over_ranks([float(rank + 1)])
        fps = shard.whole_job_throughput(frames_per_rank=48, world=world, seconds_max=t_max * 1e-3)
        out.put((rank, mine["ids"].tolist(), tuple(mine["search"].shape), t_max, fps))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_ranks_shard_and_reduce_times():
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ids = sum((r[1] for r in res), [])
    assert ids == list(range(10))                       # disjoint, complete, in order
    assert res[0][2] == (5, 4, 3) and res[1][2] == (5, 4, 3)
    assert all(r[3] == 2.0 for r in res)                # max over ranks of (1 ms, 2 ms)
    assert all(abs(r[4] - 48 * 2 / 2e-3) < 1e-6 for r in res)


def test_hot_path_net_has_exactly_the_hot_path_parameters():
    """ptt_b200.train.HotPathNet (the trainable twin used by tools/train_step.py under DDP) exposes the reference's
    parameter names for the hot path: its state_dict keys and shapes are those of synth.hot_path_layout(), which is what
    HotPath consumes and what a reference checkpoint provides.  Construction only: no CUDA needed."""
    from ptt_b200 import synth, train

    net = train.HotPathNet()
    got = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    want = {k: tuple(v) for k, v in synth.hot_path_layout().items()}
    assert got == want
    assert sum(p.numel() for p in net.parameters()) == 4106944


def test_bench_algorithmic_work_matches_the_survey_figures():
    """bench.py's per-frame algorithmic FLOPs (the numerator of roofline.achieved) are SURVEY.md 8(d)'s: SA MLPs 3.65 GFLOP +
    transformer blocks 5.24 GFLOP = 8.89 GFLOP per frame at ptt.yaml sizes."""
    import argparse
    import os
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench

    a = argparse.Namespace(nsearch=1024, ntemplate=512, batch=48)
    f = bench.algorithmic(a)
    sa = sum(v for k, v in f.items() if k.endswith(".mlp"))
    tr = sum(v for k, v in f.items() if k.endswith(".transformer"))
    assert abs(sa / 1e9 - 3.65) < 0.01 and abs(tr / 1e9 - 5.24) < 0.01 and abs((sa + tr) / 1e9 - 8.89) < 0.01
    assert set(bench.algorithmic_bytes(a)) == {"search.ball_query", "template.ball_query", "search.sa1.fps", "template.sa1.fps",
                                               "box.sa.fps", "box.sa.ball_query"}
    # one launch answers the three ball queries of a branch: B*(12N + 12M + 4*M*ns) summed over the levels
    assert bench.algorithmic_bytes(a)["search.ball_query"] == sum(
        12.0 * n + 12.0 * m + 4.0 * m * 32 for n, m in ((1024, 512), (512, 256), (256, 128)))
