"""GPU bring-up / diagnosis script (developer tool, not part of the product path).

Runs one isolated experiment per process so a trapped kernel cannot take the others down:
    python tools/bringup.py --case conv1|conv3|encoder|time ...
Each case compares the CUDA path with the CPU oracle and prints error statistics.
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

from oracle import voicemap_oracle as O  # noqa: E402
from voicemap_b200 import _lib  # noqa: E402
from voicemap_b200.engine import EncoderEngine  # noqa: E402


def relerr(a, ref):
    a = np.asarray(a, np.float64)
    ref = np.asarray(ref, np.float64)
    return float(np.abs(a - ref).max() / (np.abs(ref).max() + 1e-30)), float(
        np.linalg.norm(a - ref) / (np.linalg.norm(ref) + 1e-30))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", required=True)
    ap.add_argument("--n", type=int, default=2)
    ap.add_argument("--l", type=int, default=2048)
    ap.add_argument("--filters", type=int, default=128)
    ap.add_argument("--emb", type=int, default=64)
    ap.add_argument("--precision", type=int, default=2)
    ap.add_argument("--block", type=int, default=2)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--randbn", type=int, default=1)
    args = ap.parse_args()

    lib = _lib.load()
    params = O.init_encoder_params(args.filters, args.emb, seed=0, randomize_bn=bool(args.randbn),
                                   random_bias=True)
    eng = EncoderEngine(args.filters, args.emb, precision=args.precision)
    eng.set_weights(params)
    x = O.synthetic_clips(args.n, args.l, seed=1234)
    xd = torch.from_numpy(x[:, :, 0].copy()).cuda()

    if args.case in ("conv1", "conv3", "encoder"):
        emb64, inter64, gmax64, _ = O.encoder_forward(x, params, torch.float64, return_intermediates=True)
        emb32, inter32, gmax32, _ = O.encoder_forward(x, params, torch.float32, return_intermediates=True)

    if args.case == "conv1":
        hi, lo = eng.block1(xd)
        y = eng.merge_planes(hi, lo).cpu().numpy()
        torch.cuda.synchronize()
        print("conv1 out", y.shape, "ref", inter64[0].shape)
        print("conv1 vs fp64 (max-rel, l2-rel):", relerr(y, inter64[0]))
        print("fp32 oracle vs fp64             :", relerr(inter32[0], inter64[0]))
    elif args.case == "conv3":
        b = args.block
        src = torch.from_numpy(inter32[b - 2]).cuda()
        hi, lo = eng.split_planes(src)
        if b < 4:
            oh, ol = eng.block3(b, hi, lo)
            y = eng.merge_planes(oh, ol).cpu().numpy()
            # oracle continuation from the fp32 intermediate (so the comparison isolates this block)
            ref64 = _block_ref(inter32[b - 2], params, b, torch.float64)
            ref32 = _block_ref(inter32[b - 2], params, b, torch.float32)
            torch.cuda.synchronize()
            print(f"block{b} vs fp64:", relerr(y, ref64))
            print("fp32 oracle vs fp64:", relerr(ref32, ref64))
        else:
            part = eng.block3(4, hi, lo, gmax=True)
            emb, g = eng.gmax_dense(part, with_gmax=True)
            torch.cuda.synchronize()
            ref64 = _block_ref(inter32[2], params, 4, torch.float64).max(axis=1)
            print("block4+gmax vs fp64:", relerr(g.cpu().numpy(), ref64))
    elif args.case == "encoder":
        emb = eng.forward(xd).cpu().numpy()
        torch.cuda.synchronize()
        print("encoder emb vs fp64 :", relerr(emb, emb64))
        print("fp32 oracle vs fp64 :", relerr(emb32, emb64))
        per_clip = np.linalg.norm(emb - emb32, axis=1) / np.linalg.norm(emb32, axis=1)
        print("per-clip l2-rel vs fp32 oracle: max", per_clip.max())
    elif args.case == "time":
        for _ in range(3):
            eng.forward(xd)
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(args.iters):
            eng.forward(xd)
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / args.iters
        print(f"encoder N={args.n} L={args.l} f={args.filters} precision={args.precision}: {ms:.3f} ms/iter, "
              f"{args.n / ms * 1e3:.0f} clips/s, {args.n * 3 / ms * 1e3:.0f} audio-s/s (3 s clips)")
        # per-kernel timings
        hi, lo = eng.block1(xd)
        torch.cuda.synchronize()
        def tm(fn):
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / args.iters
        # caller-owned outputs: no allocator work between launches, the loop stays GPU-paced
        print("  block1 ms", tm(lambda: eng.block1(xd, out=(hi, lo))))
        h2, l2 = eng.block3(2, hi, lo)
        print("  block2 ms", tm(lambda: eng.block3(2, hi, lo, out=(h2, l2))))
        h3, l3 = eng.block3(3, h2, l2)
        print("  block3 ms", tm(lambda: eng.block3(3, h2, l2, out=(h3, l3))))
        part = eng.block3(4, h3, l3, gmax=True)
        print("  block4 ms", tm(lambda: eng.block3(4, h3, l3, gmax=True, out=part)))
    else:
        raise SystemExit("unknown case")


def _block_ref(x_in, params, b, dtype):
    h = O._t(x_in, dtype)
    h = O.conv1d_same_relu(h, O._t(params[f"conv{b}_kernel"], dtype), O._t(params[f"conv{b}_bias"], dtype))
    h = O.batchnorm_eval(h, O._t(params[f"bn{b}_gamma"], dtype), O._t(params[f"bn{b}_beta"], dtype),
                         O._t(params[f"bn{b}_mean"], dtype), O._t(params[f"bn{b}_var"], dtype))
    h = O.maxpool1d_valid(h, O.POOLS[b - 1])
    return h.numpy()


if __name__ == "__main__":
    main()
