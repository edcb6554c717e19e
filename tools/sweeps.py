#!/usr/bin/env python
"""BASELINE configs 4 and 5 as written (developer report, one JSON object per line, rank 0 prints):

  C4  n_seconds sweep: clips of 1 / 3 / 5 / 10 s plus the half-second lengths of the reference's own sweep
      (experiments/n_seconds_accuracy.py:23 -- 1.0 .. 6.0 s in steps of 0.5; odd lengths exercise the 'valid' pooling
      tails, e.g. 1.5 s: 6000 -> 1500 -> 750 -> 375 -> 187), entering the encoder decimated x4 as in the reference
      (voicemap/utils.py:29) and, for 1/3/5/10 s, raw 16 kHz.  The batch is AUTO-SIZED to the GPU's memory (the
      encoder's workspace + the input, --mem-fraction of the free bytes) and that batch is what runs: one
      ``vm_encoder_fwd`` call over all of it, device timed.  One GPU.
  C5  filter-width sweep 16 .. 512 (experiments/grid_search_siamese_network.py:23 sweeps 16..128; BASELINE config[4]
      names 64 -> 512) at 256 clips x 12000 samples, and k-way = 5 verification evaluation
      (voicemap/utils.py:104-216, experiments/k_way_accuracy.py:28-29) of every width on a synthetic speaker corpus:
      siamese 1-shot (pairwise head) and embedding-mode 5-shot (class means + distance + arg-min on the device,
      ``vm_nshot_score``).  Under torchrun the tasks are sharded over the ranks (every rank draws and scores its own
      tasks, the solved counts are summed) and the encoder timing is the max over ranks; ``gather_rows`` is exercised
      by embedding one fixed clip set sharded over the ranks and checking the gathered rows against rank 0's own
      embedding of the whole set (bit-identical: eval-mode embeddings do not depend on batch composition).

    python tools/sweeps.py [--configs 4,5] [--mem-fraction 0.85] [--tasks 400]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/sweeps.py --configs 5
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "examples"))
from voicemap_b200 import _lib, parallel, utils  # noqa: E402
from voicemap_b200.librispeech import LibriSpeechDataset  # noqa: E402
from voicemap_b200.models import build_siamese_net, get_baseline_convolutional_encoder  # noqa: E402

PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else dict(hbm_gbs=6650.0, bf16_tflops=1590.0)


def bound_us(length, f):
    """BASELINE.md section 3 bound per clip against the measured bf16 peak: sum_blocks max(bytes / HBM, flops / peak)."""
    tot, l, cin = 0.0, length, 1
    for i, (k, mult, pool) in enumerate(((32, 1, 4), (3, 2, 2), (3, 3, 2), (3, 4, 2))):
        cout, lout = mult * f, l // pool
        flop = 2.0 * l * k * cin * cout
        byts = 4.0 * (l * cin + (lout * cout if i < 3 else cout))
        tot += max(byts / (PEAKS["hbm_gbs"] * 1e9), flop / (PEAKS["bf16_tflops"] * 1e12))
        l, cin = lout, cout
    return tot * 1e6


def time_forward(eng, x, out, iters):
    for _ in range(2):
        eng.forward(x, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        eng.forward(x, out=out)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def config4(args, dev, rank):
    if rank != 0:
        return
    lib = _lib.load()
    enc = get_baseline_convolutional_encoder(128, 64)
    enc.precision = 2
    eng = enc._get_engine()
    eng.pack()
    ref_seconds = [1.0, 1.5, 2.0, 2.5, 3.0, 3.5, 4.0, 4.5, 5.0, 5.5, 6.0]       # n_seconds_accuracy.py:23
    plan = [(4000, s) for s in sorted(set(ref_seconds + [10.0]))] + [(16000, s) for s in (1.0, 3.0, 5.0, 10.0)]
    for rate, secs in plan:
        length = int(rate * secs)
        torch.cuda.empty_cache()
        free, total = torch.cuda.mem_get_info()
        per_clip = lib.vm_encoder_workspace_bytes(256, length, 128, 4) / 256 + 4 * length + 4 * 64
        n = int(args.mem_fraction * free / per_clip)
        n_180 = int(0.9 * 180e9 / per_clip)
        x = torch.empty((n, length), dtype=torch.float32, device=dev)
        for lo in range(0, n, 4096):                      # filled in pieces: randn of the whole batch would double it
            x[lo:lo + 4096].normal_(0.0, 0.038021)
        out = torch.empty((n, 64), dtype=torch.float32, device=dev)
        ms = time_forward(eng, x, out, iters=3)
        finite = bool(torch.isfinite(out).all().item())
        us = ms * 1e3 / n
        ls = [length]
        for p in (4, 2, 2, 2):
            ls.append(ls[-1] // p)
        print(json.dumps(dict(config="C4 n_seconds sweep, batch auto-sized to HBM", sample_rate=rate, seconds=secs,
                              length=length, pooled_lengths=ls[1:], batch=n, batch_if_180GB_free=n_180,
                              hbm_free_gb=round(free / 1e9, 1), workspace_gb=round(n * per_clip / 1e9, 1),
                              ms=round(ms, 2), audio_s_per_s=round(n * secs / (ms * 1e-3), 1),
                              us_per_clip=round(us, 3), frac_of_bf16_roofline=round(bound_us(length, 128) / us, 4),
                              finite=finite, precision=2)), flush=True)
        del x, out
        eng._workspace = None
        eng._ws_key = None


def config5(args, dev, rank, world):
    from synthetic_speakers import SyntheticCorpus
    corpus = SyntheticCorpus(20, 8, seconds=(3.2, 3.6), seed=3)
    ds = LibriSpeechDataset("synthetic", 3, stochastic=False, index=corpus.index, reader=corpus.reader)
    pre = utils.BatchPreProcessor("siamese", utils.preprocess_instances(4))
    tasks = max(1, args.tasks // world)
    g = torch.Generator().manual_seed(5)
    fixed = (0.038021 * torch.randn(64 * world, 12000, generator=g))          # the same clips on every rank
    for f in (16, 32, 64, 128, 256, 512):
        torch.manual_seed(0)
        np.random.seed(0)                                  # same initial weights on every rank (builders take no seed)
        enc = get_baseline_convolutional_encoder(f, 64, dropout=0.0)
        sia = build_siamese_net(enc, (12000, 1))
        parallel.broadcast_weights_(sia)
        eng = enc._get_engine()
        x = (0.038021 * torch.randn(256, 12000, generator=torch.Generator().manual_seed(9 + rank))).to(dev)
        out = torch.empty((256, 64), dtype=torch.float32, device=dev)
        ms = parallel.max_over_ranks(time_forward(eng, x, out, iters=20), dev)
        # sharded embedding + all-gather of the rows vs one rank embedding everything
        lo, hi = parallel.shard_bounds(fixed.shape[0])
        mine = eng.forward(fixed[lo:hi].to(dev))
        gathered = parallel.gather_rows(mine)
        whole = eng.forward(fixed.to(dev))
        gather_ok = bool(torch.equal(gathered, whole))
        res = {}
        for name, n_shot, kwargs in (("siamese_1shot_pairwise", 1, dict(network_type="siamese")),
                                     ("embedding_5shot_euclidean", 5, dict(network_type="siamese", distance="euclidean")),
                                     ("embedding_5shot_dot_product", 5, dict(network_type="siamese", distance="dot_product"))):
            np.random.seed(1000 + rank)                    # every rank draws its own tasks
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            solved = utils.n_shot_task_evaluation(sia, ds, pre, tasks, n_shot, 5, tasks_per_launch=min(tasks, 64), **kwargs)
            dt = parallel.max_over_ranks(time.perf_counter() - t0, dev)
            acc = parallel.global_mean(solved, tasks, dev)
            res[name] = dict(tasks=tasks * world, accuracy_untrained=round(acc, 4),
                             tasks_per_s_incl_host_sampling=round(tasks * world / dt, 1))
        if rank == 0:
            print(json.dumps(dict(config="C5 filters sweep + 5-way evaluation", n_gpus=world, filters=f,
                                  encoder_ms_256clips_per_gpu=round(ms, 4),
                                  audio_s_per_s=round(world * 256 * 3 / (ms * 1e-3), 1),
                                  frac_of_bf16_roofline=round(bound_us(12000, f) * 256 / (ms * 1e3), 4),
                                  sharded_embeddings_gathered_equal_single_rank=gather_ok, precision=enc.precision,
                                  five_way=res)), flush=True)
        del enc, sia, eng, x, out
        torch.cuda.empty_cache()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="4,5")
    ap.add_argument("--mem-fraction", type=float, default=0.85)
    ap.add_argument("--tasks", type=int, default=400)
    args = ap.parse_args()
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    rank, world = parallel.init_from_env("nccl", dev)
    os.environ.setdefault("TQDM_DISABLE", "1")              # progress bars off: the log is the report
    if "4" in args.configs.split(","):
        config4(args, dev, rank)
    if world > 1:
        torch.distributed.barrier()
    if "5" in args.configs.split(","):
        config5(args, dev, rank, world)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
