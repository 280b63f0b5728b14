#!/usr/bin/env python
"""Training-step timing of the hot path under DistributedDataParallel (BASELINE.json configs[3]).

    python tools/train_step.py [--batch 48] [--steps 10] [--warmup 3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/train_step.py

One rank per GPU; every rank trains on its own 48 synthetic frames per step (weak scaling).  A step = forward in train()
mode (BatchNorm batch statistics), a synthetic L2 loss on the block outputs, backward, DDP gradient all-reduce over NCCL
(the ONLY collective of the step) and an SGD update.  Timed on the device with CUDA events, max over ranks; rank 0 prints
one JSON line.  The training arithmetic is the reference's decomposition over our CUDA ops (ptt_b200/train.py)."""
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
    a = ap.parse_args()
    import torch
    import torch.distributed as dist

    from ptt_b200 import shard, synth, train

    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    net = train.HotPathNet()
    synth.load_filled(net, seed=0)
    net = net.to(dev).train()
    model = torch.nn.parallel.DistributedDataParallel(net, device_ids=[local]) if world > 1 else net
    opt = torch.optim.SGD(model.parameters(), lr=1e-4, momentum=0.9)
    n_params = sum(p.numel() for p in net.parameters())
    sets = [(torch.from_numpy(synth.make_clouds(a.batch, 1024, 3000 + 16 * rank + i, "dense")).to(dev),
             torch.from_numpy(synth.make_clouds(a.batch, 512, 4000 + 16 * rank + i, "dense", role="template")).to(dev)) for i in range(4)]

    def step(i):
        s, t = sets[i % 4]
        out = model(s, t)
        loss = sum((v.float() ** 2).mean() for v in out.values())
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss

    for i in range(a.warmup):
        step(i)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    losses = [step(i) for i in range(a.steps)]
    e1.record()
    torch.cuda.synchronize()
    ms = shard.max_over_ranks([e0.elapsed_time(e1)], device=dev)[0]
    if rank == 0:
        print(json.dumps({"metric": "training frames/sec (hot path, fwd + bwd + DDP all-reduce + SGD)",
                          "value": a.batch * world * a.steps / (ms * 1e-3), "unit": "frames/s", "n_gpus": world,
                          "ms_per_step": ms / a.steps, "steps": a.steps, "warmup": a.warmup, "scaling": "weak",
                          "batch_per_gpu": a.batch, "parameters": n_params,
                          "allreduce_bytes_per_step": 4 * n_params if world > 1 else 0,
                          "collective": "NCCL gradient all-reduce (DistributedDataParallel)" if world > 1 else "none",
                          "loss_first_last": [float(losses[0]), float(losses[-1])],
                          "mode": "train (BatchNorm batch statistics); reference decomposition over ptt_b200 CUDA ops under autograd"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
