#!/usr/bin/env python
"""Training-step timing of the hot path under DistributedDataParallel (BASELINE.json configs[3]).

    python tools/train_step.py [--batch 48] [--steps 10] [--warmup 3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/train_step.py

One rank per GPU; every rank trains on its own 48 synthetic frames per step (weak scaling).  A step = forward in train()
mode (BatchNorm batch statistics), a synthetic L2 loss on the block outputs, backward, DDP gradient all-reduce over NCCL
(the ONLY collective of the step), clip_grad_norm_(10) and an Adam update, all on the native kernels (ptt_b200/train_ops.py).
By default the whole step is captured once into a CUDA graph and the timed steps are replays (ptt_b200/train.py);
--eager times plain launches.  Timed on the device with CUDA events, max over ranks; rank 0 prints one JSON line."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=48)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--eager", action="store_true", help="time eager launches instead of CUDA-graph replays of the step")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist

    from ptt_b200 import shard, train

    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        train.prepare_env_for_graphs()
        dist.init_process_group("nccl", device_id=dev)
    r = train.time_train_step(dev, world, rank, local, batch=a.batch, steps=a.steps, warmup=a.warmup, graph=not a.eager)
    keys = [k for k in ("ms_per_step", "host_enqueue_ms_per_step") if k in r]
    for k, v in zip(keys, shard.max_over_ranks([r[k] for k in keys], device=dev)):   # slowest rank
        r[k] = v
    ms = r["ms_per_step"]
    if rank == 0:
        print(json.dumps({"metric": "training frames/sec (hot path, fwd + bwd + DDP all-reduce + clip + Adam)",
                          "value": a.batch * world / (ms * 1e-3), "unit": "frames/s", "n_gpus": world, **r, "ms_per_step": ms,
                          "scaling": "weak",
                          "collective": "NCCL gradient all-reduce (DistributedDataParallel)" if world > 1 else "none"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
