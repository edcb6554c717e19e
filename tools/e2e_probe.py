"""Where does end-to-end predict() time go?  Times model.predict on a pinned host batch for several chunk plans
(voicemap_b200.models._pipeline_plan), plus the raw H2D copy and the device-only forward, so the pipeline plan
can be chosen from measurements.  Run on a GPU box: python tools/e2e_probe.py"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import voicemap_oracle as O  # noqa: E402  (weights/inputs only)
from voicemap_b200 import models  # noqa: E402

N, L = 256, 12000


def wall(fn, iters=40, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / iters * 1e3


def main():
    params = O.init_encoder_params(128, 64, seed=0, randomize_bn=True, random_bias=True)
    enc = models.get_baseline_convolutional_encoder(128, 64, dropout=0.0)
    enc.set_named_weights(params)
    g = torch.Generator().manual_seed(3)
    sets = [(O.WHITEN_RMS * torch.randn(N, L, 1, generator=g)).pin_memory() for _ in range(6)]
    dev = [s[:, :, 0].cuda().contiguous() for s in sets]
    eng = enc._get_engine()
    out = torch.empty((N, 64), device="cuda")
    k = [0]

    def fwd():
        k[0] += 1
        eng.forward(dev[k[0] % 6], out=out)

    def h2d():
        k[0] += 1
        dev[0].copy_(sets[k[0] % 6][:, :, 0], non_blocking=True)

    print(f"device forward      {wall(fwd):.3f} ms")
    t = wall(h2d)
    print(f"H2D 12.3 MB pinned  {t:.3f} ms  ({N * L * 4 / t / 1e6:.1f} GB/s)")
    def zc():
        k[0] += 1
        eng.forward(sets[k[0] % 6][:, :, 0], out=out)
    print(f"zero-copy forward (block 1 reads pinned host memory)  {wall(zc):.3f} ms")

    def zc_e2e():
        k[0] += 1
        eng.forward(sets[k[0] % 6][:, :, 0], out=out)
        return out.cpu().numpy()
    print(f"zero-copy forward + D2H                               {wall(zc_e2e):.3f} ms")

    def split_e2e(frac):
        # first part zero-copy, rest by DMA issued up front on the side stream
        k[0] += 1
        x = sets[k[0] % 6][:, :, 0]
        m = int(N * frac) // 8 * 8
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            dev[1][m:].copy_(x[m:], non_blocking=True)
            evc = torch.cuda.Event(); evc.record(side)
        eng.forward(x[:m], out=out[:m])
        torch.cuda.current_stream().wait_event(evc)
        eng.forward(dev[1][m:], out=out[m:])
        return out.cpu().numpy()
    side = torch.cuda.Stream()
    for frac in (0.25, 0.5, 0.75):
        print(f"zero-copy first {frac:.2f} + DMA rest + D2H                  {wall(lambda: split_e2e(frac)):.3f} ms")
    plans = {
        "single": [N],
        "2 equal": [128, 128],
        "default": None,
        "16/48/192": [16, 48, 192],
        "32/96/128": [32, 96, 128],
        "8/24/64/160": [8, 24, 64, 160],
        "64/192": [64, 192],
        "32/224": [32, 224],
    }
    orig = models._pipeline_plan
    for name, sizes in plans.items():
        if sizes is None:
            models._pipeline_plan = orig
        else:
            def plan(n, pinned=True, sizes=sizes):
                lo, out_ = 0, []
                for s in sizes:
                    out_.append((lo, lo + s))
                    lo += s
                return out_
            models._pipeline_plan = plan

        def pred():
            k[0] += 1
            enc.predict(sets[k[0] % 6])
        ms = wall(pred)
        print(f"predict plan {name:12s} {ms:.3f} ms  -> {N * 3000 / ms:.0f} audio-s/s")
    models._pipeline_plan = orig


if __name__ == "__main__":
    main()
