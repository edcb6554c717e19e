#!/usr/bin/env python
"""CPU numerics study (no GPU): embedding error of split-precision schemes for the k=3 convolutions (blocks 2-4).

Every scheme computes  acc = Xh*Wh + C1 + C2  with Xh = fp16(X), Wh = fp16(W) and different roundings of the two
correction products C1 ~ Xl*Wh, C2 ~ Xh*Wl (Xl = X - Xh, Wl = W - Wh).  Products and sums are evaluated in float64, so
the table isolates operand rounding (the tensor core's fp32 accumulation adds ~5e-6, measured on the GPU).  Block 1
(K = 32, Cin = 1) always uses the shipped fp16 x 3 scheme.

    python tools/sim_split_precision.py [--clips 8] [--length 12000] [--stress 1]
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import voicemap_oracle as O  # noqa: E402

F64 = torch.float64


def q_fp16(t):
    return t.to(torch.float32).to(torch.float16).to(F64)


def q_f8(t, kind, scale):
    """Round t * 2**scale to e4m3 / e5m2 (saturating), return the value / 2**scale."""
    dt, top = (torch.float8_e4m3fn, 448.0) if kind == "e4m3" else (torch.float8_e5m2, 57344.0)
    s = float(2.0 ** scale)
    v = (t * s).clamp(-top, top).to(torch.float32).to(dt).to(F64)
    return v / s


def conv_same(x, w):
    """x (N, L, Cin), w (3, Cin, Cout) float64 -> (N, L, Cout), cross-correlation with 1/1 zero padding."""
    y = F.conv1d(x.transpose(1, 2), w.permute(2, 1, 0), padding=1)
    return y.transpose(1, 2)


def run(params, x, scheme):
    h = torch.from_numpy(x).to(F64)
    for i in range(1, 5):
        w = torch.from_numpy(params[f"conv{i}_kernel"]).to(F64)
        b = torch.from_numpy(params[f"conv{i}_bias"]).to(F64)
        h32 = h.to(torch.float32).to(F64)
        xh, wh = q_fp16(h32), q_fp16(w)
        xl, wl = h32 - xh, w - wh
        if i == 1:
            pad = F.pad(h32.transpose(1, 2), (15, 16))
            padh, padl = q_fp16(pad), q_fp16(pad - q_fp16(pad))
            wk = w.permute(2, 1, 0)
            wkh, wkl = q_fp16(wk), q_fp16(wk - q_fp16(wk))
            acc = (F.conv1d(padh, wkh) + F.conv1d(padl, wkh) + F.conv1d(padh, wkl)).transpose(1, 2)
        elif scheme == "fp16x1":
            acc = conv_same(xh, wh)
        elif scheme == "fp16x3":
            acc = conv_same(xh, wh) + conv_same(q_fp16(xl), wh) + conv_same(xh, q_fp16(wl))
        else:
            kx, kw, a, bsc = scheme          # formats of the X-side / W-side fp8 operands, scales of Xl and Wl
            c1 = conv_same(q_f8(xl, kx, a), q_f8(wh, kw, -a if kx == kw == "e5m2" else 4))
            c2 = conv_same(q_f8(xh, kx, -bsc if kx == kw == "e5m2" else 0), q_f8(wl, kw, bsc))
            acc = conv_same(xh, wh) + c1 + c2
        u = torch.relu(acc + b)
        s = torch.from_numpy(params[f"bn{i}_gamma"]).to(F64) / torch.sqrt(
            torch.from_numpy(params[f"bn{i}_var"]).to(F64) + O.BN_EPS)
        t = torch.from_numpy(params[f"bn{i}_beta"]).to(F64) - torch.from_numpy(params[f"bn{i}_mean"]).to(F64) * s
        y = u * s + t
        p = O.POOLS[i - 1]
        lo = y.shape[1] // p
        h = y[:, :lo * p].reshape(y.shape[0], lo, p, y.shape[2]).amax(dim=2)
    g = h.amax(dim=1)
    return (g @ torch.from_numpy(params["dense_kernel"]).to(F64) + torch.from_numpy(params["dense_bias"]).to(F64)).numpy()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=8)
    ap.add_argument("--length", type=int, default=12000)
    ap.add_argument("--filters", type=int, default=128)
    ap.add_argument("--stress", type=int, default=1, help="1: randomised BatchNorm (parity stress), 0: Keras init")
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    params = O.init_encoder_params(args.filters, 64, seed=args.seed, randomize_bn=bool(args.stress), random_bias=True)
    x = O.synthetic_clips(args.clips, args.length, seed=1234)
    ref = O.encoder_forward(x, params, torch.float64)
    schemes = [("fp16x1", "fp16x1"), ("fp16x3 (shipped)", "fp16x3"),
               ("fp16 + e4m3 x e4m3 corrections, free scales (second accumulator / MX scale)", ("e4m3", "e4m3", 12, 16)),
               ("fp16 + e5m2 x e5m2 corrections, shared accumulator (scales 2^6 / 2^-6, 2^-4 / 2^4)", ("e5m2", "e5m2", 6, 4)),
               ("fp16 + e5m2 x e5m2, scales 2^8 / 2^-8, 2^-2 / 2^2", ("e5m2", "e5m2", 8, 2))]
    for name, sch in schemes:
        emb = run(params, x, sch)
        rel = np.linalg.norm(emb - ref, axis=1) / np.linalg.norm(ref, axis=1)
        mx = np.abs(emb - ref).max(axis=1) / np.abs(ref).max(axis=1)
        print(f"{name:90s} per-clip l2-rel max {rel.max():.2e} mean {rel.mean():.2e}   max-abs/max-ref {mx.max():.2e}")


if __name__ == "__main__":
    main()
