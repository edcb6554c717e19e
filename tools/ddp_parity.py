#!/usr/bin/env python
"""Multi-GPU parity of the data-parallel train step: W ranks, each holding 1/W of the pairs, with
synchronised BatchNorm and one flat gradient all-reduce, must reproduce the single-device step on the whole batch
(what the reference's Keras does): loss, every gradient, BN moving statistics, and the weights after one Adam update.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/ddp_parity.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from voicemap_b200 import parallel  # noqa: E402
from voicemap_b200.keras_compat import Adam  # noqa: E402
from voicemap_b200.models import build_siamese_net, get_baseline_convolutional_encoder  # noqa: E402
from voicemap_b200.training import TrainEngine  # noqa: E402

PAIRS, LENGTH, FILTERS, EMB = 16, 4000, 64, 32


def build(seed):
    from voicemap_b200.models import EncoderModel
    enc = EncoderModel(FILTERS, EMB, dropout=0.0, seed=seed)
    rng = np.random.default_rng(seed + 1)
    named = {}
    for i in range(1, 5):
        named[f"bn{i}_gamma"] = rng.uniform(0.5, 1.5, enc.weights[f"bn{i}_gamma"].shape).astype(np.float32)
        named[f"bn{i}_beta"] = rng.normal(0, 0.2, enc.weights[f"bn{i}_beta"].shape).astype(np.float32)
    enc.set_named_weights(named)
    sia = build_siamese_net(enc, (LENGTH, 1))
    sia.head_weights["head_kernel"][:] = 0.05
    sia.head_weights["head_bias"][:] = -0.3
    opt = Adam(clipnorm=1.0)
    sia.compile(loss="binary_crossentropy", optimizer=opt)
    return sia, TrainEngine(sia, opt, sia.loss)


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def main():
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    rank, world = parallel.init_from_env("nccl", dev)
    g = torch.Generator().manual_seed(7)
    x1 = (0.038021 * torch.randn(PAIRS, LENGTH, generator=g)).numpy()
    x2 = (0.038021 * torch.randn(PAIRS, LENGTH, generator=g)).numpy()
    y = (np.arange(PAIRS) % 2).astype(np.float32)
    lo, hi = parallel.shard_bounds(PAIRS, rank, world)

    from voicemap_b200.training import sync_bn_peers
    results = {}
    # "p2p": sums cross the ranks inside the kernels (NVLink peer memory) + bucketed gradient all-reduce (the default of
    # fit_generator); "nccl": 8 small all-reduce calls + one flat gradient all-reduce; False: per-rank statistics
    for sync in ("p2p", "nccl", False):
        sia, tr = build(0)
        allreduce = parallel.allreduce_sum_ if world > 1 else None
        peers = sync_bn_peers() if (sync == "p2p" and world > 1) else None
        tr.set_sync_bn(allreduce if sync else None, world, peers=peers)
        tr.set_gradient_buckets(sync == "p2p" and world > 1)
        if sync == "p2p" and world > 1:    # a few steps on a throw-away model first: the exchange buffers' two slots and
            _, warm = build(1)             # the sequence counter are then mid-stream when the compared step runs
            warm.set_sync_bn(allreduce, world, peers=peers)
            warm.set_gradient_buckets(True)
            for _ in range(3):
                warm.siamese_step(x1[lo:hi], x2[lo:hi], y[lo:hi], apply=True, allreduce=allreduce, world=world)
            del warm
        lv, _ = tr.siamese_step(x1[lo:hi], x2[lo:hi], y[lo:hi], apply=True, allreduce=allreduce, world=world)
        loss = parallel.global_mean(float(lv.item()) * (hi - lo), hi - lo, dev)
        grads = {k: v / world for k, v in tr.gradients().items()}
        results[sync] = (loss, grads, {k: v.cpu().numpy() for k, v in tr.moving.items()},
                         {k: v.detach().cpu().numpy().copy() for k, v in tr.p.items()})
    if rank == 0:
        sia, tr = build(0)
        lv, _ = tr.siamese_step(x1, x2, y, apply=True)
        ref = (float(lv.item()), tr.gradients(), {k: v.cpu().numpy() for k, v in tr.moving.items()},
               {k: v.detach().cpu().numpy().copy() for k, v in tr.p.items()})
        ok = True
        for sync in ("p2p", "nccl", False):
            loss, grads, moving, params = results[sync]
            floor = 1e-3 * max(np.abs(v).max() for v in ref[1].values())
            gerr = max(np.abs(grads[k] - ref[1][k]).max() / max(np.abs(ref[1][k]).max(), floor) for k in ref[1])
            merr = max(rel(moving[k], ref[2][k]) for k in ref[2])
            perr = max(np.abs(params[k] - ref[3][k]).max() for k in ref[3])
            print(f"world {world} sync_bn={sync}: loss {loss:.7f} vs single-device {ref[0]:.7f} "
                  f"(rel {abs(loss - ref[0]) / abs(ref[0]):.2e}); worst gradient tensor rel err {gerr:.2e}; "
                  f"moving statistics rel err {merr:.2e}; max |weight diff| after Adam {perr:.2e}", flush=True)
            if sync:
                ok = ok and abs(loss - ref[0]) <= 1e-5 * abs(ref[0]) and gerr <= 1e-4 and merr <= 1e-5
        print("DDP PARITY", "OK" if ok else "FAILED", flush=True)
        if not ok:
            sys.exit(1)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
