#!/usr/bin/env python
"""BASELINE configs 4 and 5 on one GPU (developer report, JSON lines):
  C4  n_seconds sweep (the reference's n_seconds_accuracy.py lengths): clips of 1/3/5/10 s entering the encoder
      decimated (4 kHz: L = 4000..40000) and raw (16 kHz: L = 16000..160000); batch auto-sized from free HBM
      (reported) and capped for the timed run.
  C5  filters sweep {16, 32, 64, 128} (grid_search_siamese_network.py:23): encoder throughput and batched 5-way
      1-shot evaluation rate on a synthetic speaker corpus.
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "examples"))
from voicemap_b200 import _lib, utils  # noqa: E402
from voicemap_b200.librispeech import LibriSpeechDataset  # noqa: E402
from voicemap_b200.models import build_siamese_net, get_baseline_convolutional_encoder  # noqa: E402

PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else dict(hbm_gbs=6650.0, bf16_tflops=1590.0)


def bound_us(length, f):
    """BASELINE.md bound per clip: sum_blocks max(bytes / HBM, flops / (bf16 peak / 2))."""
    tot, l, cin = 0.0, length, 1
    for i, (k, mult, pool) in enumerate(((32, 1, 4), (3, 2, 2), (3, 3, 2), (3, 4, 2))):
        cout, lout = mult * f, l // pool
        flop = 2.0 * l * k * cin * cout
        byts = 4.0 * (l * cin + (lout * cout if i < 3 else cout))
        tot += max(byts / (PEAKS["hbm_gbs"] * 1e9), flop / (PEAKS["bf16_tflops"] / 2 * 1e12))
        l, cin = lout, cout
    return tot * 1e6


def time_forward(eng, x, iters=5):
    for _ in range(3):
        eng.forward(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        eng.forward(x)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    free, total = torch.cuda.mem_get_info()
    # ---- C4
    enc = get_baseline_convolutional_encoder(128, 64)
    eng = enc._get_engine()
    for rate, lengths in ((4000, (4000, 12000, 20000, 40000)), (16000, (16000, 48000, 80000, 160000))):
        for length in lengths:
            per_clip = lib.vm_encoder_workspace_bytes(64, length, 128, 4) / 64 + 4 * length
            n_auto = int(0.9 * total / per_clip)
            n = min(n_auto, max(64, int(2 ** 31 // (length * 512 * 4)) // 4, 1))      # keep index math in int32 range
            n = min(n, 2048 if length <= 20000 else 512)
            x = 0.038 * torch.randn(n, length, device=dev)
            ms = time_forward(eng, x)
            secs = length / rate
            us = ms * 1e3 / n
            print(json.dumps(dict(config="C4 n_seconds sweep", sample_rate=rate, seconds=secs, length=length,
                                  batch_timed=n, batch_auto_180GB=n_auto, ms=round(ms, 3),
                                  audio_s_per_s=round(n * secs / (ms * 1e-3), 1), us_per_clip=round(us, 3),
                                  frac_of_tf32_roofline=round(bound_us(length, 128) / us, 4))), flush=True)
            del x
            eng._workspace = None
            eng._ws_key = None
            torch.cuda.empty_cache()
    # ---- C5
    from synthetic_speakers import SyntheticCorpus
    corpus = SyntheticCorpus(20, 5, seconds=(3.2, 3.6), seed=3)
    ds = LibriSpeechDataset("synthetic", 3, stochastic=False, index=corpus.index, reader=corpus.reader)
    pre = utils.BatchPreProcessor("siamese", utils.preprocess_instances(4))
    for f in (16, 32, 64, 128):
        enc = get_baseline_convolutional_encoder(f, 64, dropout=0.0)
        sia = build_siamese_net(enc, (12000, 1))
        eng = enc._get_engine()
        x = 0.038 * torch.randn(256, 12000, device=dev)
        ms = time_forward(eng, x)
        np.random.seed(0)
        t0 = time.perf_counter()
        ok = utils.n_shot_task_evaluation_batched(sia, ds, pre, 100, 1, 5, tasks_per_launch=50)
        dt = time.perf_counter() - t0
        print(json.dumps(dict(config="C5 filters sweep", filters=f, encoder_ms_256clips=round(ms, 3),
                              audio_s_per_s=round(256 * 3 / (ms * 1e-3), 1),
                              frac_of_tf32_roofline=round(bound_us(12000, f) * 256 / (ms * 1e3), 4),
                              five_way_one_shot_tasks=100, solved_untrained=ok,
                              tasks_per_s_incl_host_sampling=round(100 / dt, 1))), flush=True)


if __name__ == "__main__":
    main()
