#!/usr/bin/env python
"""Training-step benchmark (BASELINE config[2]): siamese encoder + contrastive loss, 3 s pairs, global batch 128
pairs sharded over the ranks (16 pairs/GPU at 8 GPUs), one flat NCCL gradient all-reduce per step.

    python tools/train_bench.py [--pairs-per-gpu 16] [--steps 20]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/train_bench.py

Prints one JSON line (rank 0): pairs/s and audio-s/s of the full train step (forward, backward, all-reduce, Adam),
max over ranks, CUDA-event timed."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from voicemap_b200 import parallel  # noqa: E402
from voicemap_b200.keras_compat import Adam  # noqa: E402
from voicemap_b200.models import build_siamese_net, get_baseline_convolutional_encoder  # noqa: E402
from voicemap_b200.training import TrainEngine  # noqa: E402
from voicemap_b200.utils import contrastive_loss  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs-per-gpu", type=int, default=None)
    ap.add_argument("--global-pairs", type=int, default=128)
    ap.add_argument("--length", type=int, default=12000)
    ap.add_argument("--filters", type=int, default=128)
    ap.add_argument("--emb", type=int, default=64)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--bwd-precision", type=int, default=3)
    ap.add_argument("--local-bn", action="store_true", help="per-rank BatchNorm statistics instead of synchronised")
    ap.add_argument("--nccl-bn", action="store_true", help="BatchNorm sums through NCCL all-reduce calls instead of the "
                                                           "peer-memory kernels")
    ap.add_argument("--flat-allreduce", action="store_true", help="one gradient all-reduce after backward instead of "
                                                                  "per-block buckets inside it")
    ap.add_argument("--timeline", action="store_true", help="print live per-launch durations of a step (one rank)")
    args = ap.parse_args()
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    rank, world = parallel.init_from_env("nccl", dev)
    pairs = args.pairs_per_gpu or max(1, args.global_pairs // world)

    enc = get_baseline_convolutional_encoder(args.filters, args.emb, dropout=0.0)
    sia = build_siamese_net(enc, (args.length, 1))
    opt = Adam(clipnorm=1.0)
    sia.compile(loss=contrastive_loss, optimizer=opt)
    tr = TrainEngine(sia, opt, sia.loss, bwd_precision=args.bwd_precision)
    g = torch.Generator().manual_seed(100 + rank)
    x1 = (0.038021 * torch.randn(pairs, args.length, generator=g)).to(dev)
    x2 = (0.038021 * torch.randn(pairs, args.length, generator=g)).to(dev)
    y = np.concatenate([np.zeros(pairs // 2), np.ones(pairs - pairs // 2)]).astype(np.float32)
    allreduce = parallel.allreduce_sum_ if world > 1 else None
    from voicemap_b200.training import sync_bn_peers
    peers = sync_bn_peers() if (world > 1 and not args.local_bn and not args.nccl_bn) else None
    tr.set_sync_bn(None if args.local_bn else allreduce, world, peers=peers)
    tr.set_gradient_buckets(world > 1 and not args.flat_allreduce)

    for _ in range(max(args.warmup, 3)):
        tr.siamese_step(x1, x2, y, allreduce=allreduce, world=world)
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    import time
    e0.record()
    for _ in range(args.steps):
        lv, _ = tr.siamese_step(x1, x2, y, allreduce=allreduce, world=world)
    e1.record()
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    ms = parallel.max_over_ranks(e0.elapsed_time(e1) / args.steps, dev)
    # host time to ISSUE a step: four steps from an idle device, so the launch queue never fills and blocks the host
    t_host = time.perf_counter()
    for _ in range(4):
        tr.siamese_step(x1, x2, y, allreduce=allreduce, world=world)
    t_host = (time.perf_counter() - t_host) / 4 * 1e3
    torch.cuda.synchronize()
    if args.timeline and world == 1:
        total = 0.0
        for what, t in tr.time_siamese_step(x1, x2, y):
            total += t
            print(f"{t * 1e3:9.1f} us  {what}")
        print(f"{total * 1e3:9.1f} us  sum (events between launches; step above {ms * 1e3:.1f} us)")
    if rank == 0:
        print(json.dumps(dict(metric="siamese_train_pairs_per_sec", value=round(world * pairs / (ms * 1e-3), 1),
                              unit="pairs/s", audio_seconds_per_sec=round(world * pairs * 2 * 3.0 / (ms * 1e-3), 1),
                              n_gpus=world, ms_per_step=round(ms, 3), host_issue_ms_per_step=round(t_host, 3), steps=args.steps, scaling="strong" if
                              args.pairs_per_gpu is None else "weak",
                              config=dict(workload=f"siamese train step (fwd+bwd+Adam), contrastive loss, "
                                                   f"{pairs} pairs/GPU x {args.length} samples, filters={args.filters}, "
                                                   f"fwd fp16x3, bwd fp16 mode {args.bwd_precision}",
                                          parallelism=f"dp{world}, one flat fp32 gradient all-reduce "
                                                      f"({tr.nparams * 4 / 1e6:.1f} MB) per step, BatchNorm "
                                                      f"{'per rank' if (args.local_bn or world == 1) else ('synchronised: 8 NCCL all-reduces' if peers is None else 'synchronised: sums cross the ranks inside the kernels (NVLink peer memory)')}, gradients {'one flat all-reduce' if args.flat_allreduce else '5 buckets overlapped with backward'}"),
                              final_loss=float(lv.item()))), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
