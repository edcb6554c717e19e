"""GPU/CPU timeline of one pipelined predict(): when each chunk's copy and kernels start/finish on the device
(CUDA events) and when the host issued them (perf_counter), to see what the end-to-end path is bound by."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import voicemap_oracle as O  # noqa: E402
from voicemap_b200 import models  # noqa: E402

N, L = 256, 12000


def main():
    params = O.init_encoder_params(128, 64, seed=0, randomize_bn=True, random_bias=True)
    enc = models.get_baseline_convolutional_encoder(128, 64, dropout=0.0)
    enc.set_named_weights(params)
    g = torch.Generator().manual_seed(3)
    x = (O.WHITEN_RMS * torch.randn(N, L, 1, generator=g)).pin_memory()
    for _ in range(5):
        enc.predict(x)
    eng = enc._get_engine()
    xt = enc._host_batch(x)
    dev = eng.device
    out = torch.empty((N, 64), dtype=torch.float32, device=dev)
    xin = torch.empty((N, L), dtype=torch.float32, device=dev)
    copy_stream = torch.cuda.Stream(device=dev)
    main_s = torch.cuda.current_stream(dev)
    for rep in range(3):
        torch.cuda.synchronize()
        ev = lambda: torch.cuda.Event(enable_timing=True)
        t0e = ev(); t0e.record(main_s)
        copy_stream.wait_stream(main_s)
        t0 = time.perf_counter()
        rows = []
        for lo, hi in models._pipeline_plan(N, True):
            c0, c1, k0, k1 = ev(), ev(), ev(), ev()
            ta = time.perf_counter()
            with torch.cuda.stream(copy_stream):
                c0.record(copy_stream)
                xin[lo:hi].copy_(xt[lo:hi], non_blocking=True)
                c1.record(copy_stream)
            tb = time.perf_counter()
            main_s.wait_event(c1)
            k0.record(main_s)
            eng.forward(xin[lo:hi], out=out[lo:hi])
            k1.record(main_s)
            tc = time.perf_counter()
            rows.append((lo, hi, c0, c1, k0, k1, ta - t0, tb - t0, tc - t0))
        td = time.perf_counter()
        host = out.cpu()
        te = time.perf_counter()
        torch.cuda.synchronize()
        print(f"rep {rep}: host total {1e3 * (te - t0):.3f} ms (issue done at {1e3 * (td - t0):.3f} ms)")
        for lo, hi, c0, c1, k0, k1, ta, tb, tc in rows:
            print(f"  chunk [{lo:3d},{hi:3d})  copy dev {t0e.elapsed_time(c0):.3f}-{t0e.elapsed_time(c1):.3f} ms   "
                  f"kernels dev {t0e.elapsed_time(k0):.3f}-{t0e.elapsed_time(k1):.3f} ms   "
                  f"host issue copy@{1e3 * ta:.3f} fwd@{1e3 * tb:.3f} done@{1e3 * tc:.3f}")


if __name__ == "__main__":
    main()
