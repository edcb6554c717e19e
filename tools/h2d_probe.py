#!/usr/bin/env python
"""Developer probe: host->device copy rate of a pinned batch alone and while the encoder kernels run on another stream
(the e2e path's copy/compute overlap, profiles/r01_e2e_probe.log)."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import voicemap_oracle as O  # noqa: E402
from voicemap_b200.engine import EncoderEngine  # noqa: E402


def main():
    n, length, reps = 256, 12000, 20
    eng = EncoderEngine(128, 64)
    eng.set_weights(O.init_encoder_params(128, 64, seed=0, randomize_bn=True, random_bias=True))
    host = (0.038 * torch.randn(n, length)).pin_memory()
    xdev = host.cuda()
    dst = torch.empty_like(xdev)
    side = torch.cuda.Stream()
    main_s = torch.cuda.current_stream()
    for _ in range(3):
        eng.forward(xdev)
    torch.cuda.synchronize()

    def timed_copies(stream, k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(k):
                dst.copy_(host, non_blocking=True)
            e1.record(stream)
        return e0, e1

    e0, e1 = timed_copies(side, reps)
    torch.cuda.synchronize()
    alone = e0.elapsed_time(e1) / reps
    print(f"H2D alone            {alone:.3f} ms/copy  {host.numel() * 4 / alone / 1e6:.1f} GB/s")

    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record(main_s)
    for _ in range(reps):
        eng.forward(xdev)
    k1.record(main_s)
    torch.cuda.synchronize()
    print(f"forward alone        {k0.elapsed_time(k1) / reps:.3f} ms")

    k0.record(main_s)
    for _ in range(reps):
        eng.forward(xdev)
    k1.record(main_s)
    e0, e1 = timed_copies(side, reps)
    torch.cuda.synchronize()
    both = e0.elapsed_time(e1) / reps
    print(f"H2D under kernels    {both:.3f} ms/copy  {host.numel() * 4 / both / 1e6:.1f} GB/s")
    print(f"forward under copies {k0.elapsed_time(k1) / reps:.3f} ms")


if __name__ == "__main__":
    main()
