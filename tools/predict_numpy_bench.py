"""predict() called the way the reference calls it -- with a numpy float64 batch (voicemap/utils.py:133,156) -- next to
the same batch as a pinned float32 tensor (bench.py's e2e arm) and to the cost of the host-side cast alone."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import voicemap_oracle as O  # noqa: E402  (seeded weights only)
from voicemap_b200.models import get_baseline_convolutional_encoder  # noqa: E402


def timed(fn, reps=20):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


def main():
    n, length = 256, 12000
    enc = get_baseline_convolutional_encoder(128, 64, dropout=0.0)
    enc.set_named_weights(O.init_encoder_params(128, 64, seed=0, randomize_bn=True, random_bias=True))
    x64 = O.WHITEN_RMS * np.random.default_rng(0).standard_normal((n, length, 1))
    x32 = torch.from_numpy(x64.astype(np.float32)).pin_memory()
    a = enc.predict(x64)
    b = enc.predict(x32)
    print("bit identical:", bool(np.array_equal(a, b)))
    print("host cast alone (one thread, astype)       %.2f ms" % timed(lambda: np.ascontiguousarray(x64[:, :, 0], dtype=np.float32), 5))
    print("predict(pinned float32 tensor)             %.3f ms" % timed(lambda: enc.predict(x32)))
    print("predict(numpy float64), staged             %.3f ms  (%d cast threads)" % (timed(lambda: enc.predict(x64)), enc._host_stage.workers))
    # same-process A/B of staging orders (alternated): chunk plans x threads per chunk
    from concurrent.futures import ThreadPoolExecutor
    from voicemap_b200 import models as M
    hs = enc._host_stage
    keep = (M._cast_plan, hs.pool, hs.workers, hs.pieces)

    def fixed(sizes):
        def plan(n):
            out, lo = [], 0
            for sz in sizes:
                if lo >= n:
                    break
                out.append((lo, min(n, lo + sz)))
                lo += sz
            while lo < n:
                out.append((lo, min(n, lo + sizes[-1])))
                lo += sizes[-1]
            return out
        return plan
    plans = {"4 x 64": fixed([64]), "2 x 128": fixed([128]), "1 x 256": fixed([256]), "64 + 192": fixed([64, 192]),
             "96 + 160": fixed([96, 160]), "3 x 86": fixed([86])}
    pools = {w: ThreadPoolExecutor(max_workers=w) for w in (2, 3, 4, 6)}
    for rnd in range(2):
        for w, pool in pools.items():
            for name, plan in plans.items():
                hs.pool, hs.workers, hs.pieces = pool, w, True
                M._cast_plan = plan
                ok = bool(np.array_equal(enc.predict(x64), a))
                print("round %d  row blocks on %d threads, plan %-9s %.3f ms  bit identical %s" % (
                    rnd, w, name, timed(lambda: enc.predict(x64)), ok))
        hs.pool, hs.workers, hs.pieces = keep[1], keep[2], False
        M._cast_plan = fixed([64])
        print("round %d  one thread per chunk (4 threads), plan 4 x 64  %.3f ms" % (rnd, timed(lambda: enc.predict(x64))))
    M._cast_plan, hs.pool, hs.workers, hs.pieces = keep
    print("predict(numpy float64), staged (defaults)  %.3f ms  (%d cast threads)" % (timed(lambda: enc.predict(x64)), hs.workers))
    big = O.WHITEN_RMS * np.random.default_rng(1).standard_normal((2048, length, 1))
    print("predict(numpy float64, 2048 clips)         %.3f ms" % timed(lambda: enc.predict(big), 5))
    small = x64[:5]
    print("predict(numpy float64, 5 clips)            %.3f ms" % timed(lambda: enc.predict(small)))


if __name__ == "__main__":
    main()
