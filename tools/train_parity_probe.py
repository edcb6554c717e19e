#!/usr/bin/env python
"""Gradient / loss errors of one siamese train step against the fp64 autograd oracle, for each (forward precision,
backward precision) pair -- the measurement behind the tolerances of tests/test_gpu_train.py and the default of
TrainEngine.  python tools/train_parity_probe.py [--pairs 4 --length 1024 --filters 128 --emb 64]"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import voicemap_oracle as O  # noqa: E402
from voicemap_b200.keras_compat import Adam  # noqa: E402
from voicemap_b200.models import build_siamese_net, get_baseline_convolutional_encoder  # noqa: E402
from voicemap_b200.training import TrainEngine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=4)
    ap.add_argument("--length", type=int, default=1024)
    ap.add_argument("--filters", type=int, default=128)
    ap.add_argument("--emb", type=int, default=64)
    ap.add_argument("--loss", default="contrastive_loss")
    ap.add_argument("--modes", default="3:3,3:2,2:3,2:2,3:1")
    args = ap.parse_args()
    n, length = args.pairs, args.length
    params = O.init_encoder_params(args.filters, args.emb, seed=0, randomize_bn=False, random_bias=True)
    rng = np.random.default_rng(1)
    for i in range(1, 5):
        params[f"bn{i}_gamma"] = rng.uniform(-1.2, 1.5, params[f"bn{i}_gamma"].shape).astype(np.float32)
        params[f"bn{i}_beta"] = rng.normal(0, 0.2, params[f"bn{i}_beta"].shape).astype(np.float32)
    x1, x2 = O.synthetic_clips(n, length, seed=11), O.synthetic_clips(n, length, seed=12)
    y = (np.arange(n) >= n // 2).astype(np.float32)
    ref = None
    for mode in args.modes.split(","):
        fp, bp = (int(v) for v in mode.split(":"))
        enc = get_baseline_convolutional_encoder(args.filters, args.emb, dropout=0.0)
        enc.set_named_weights(params)
        sia = build_siamese_net(enc, (length, 1))
        sia.head_weights["head_kernel"][:] = 0.05
        sia.head_weights["head_bias"][:] = -0.3
        opt = Adam(clipnorm=1.0)
        sia.compile(loss=args.loss, optimizer=opt)
        tr = TrainEngine(sia, opt, sia.loss, precision=fp, bwd_precision=bp)
        hw, hb = sia.head_weights["head_kernel"].reshape(-1).copy(), sia.head_weights["head_bias"].copy()
        lv, _ = tr.siamese_step(x1, x2, y, apply=False)
        torch.cuda.synchronize()
        masks = [[tr.relu_pattern(b)[br * n:(br + 1) * n].cpu().numpy().astype(np.float64) for b in range(4)]
                 for br in range(2)]
        ref = O.siamese_train_step_grads(params, hw, hb, x1, x2, y, loss=args.loss, relu_masks=masks, pool_selects=[[tr.argmax_flags(b)[br * n:(br + 1) * n].cpu().numpy() for b in range(4)] for br in range(2)], gmax_selects=[tr.jstar[br * n:(br + 1) * n].cpu().numpy() for br in range(2)])
        refg = dict(ref["grads"], head_kernel=ref["head_w_grad"], head_bias=ref["head_b_grad"])
        floor = 1e-3 * max(np.abs(np.asarray(g)).max() for g in refg.values())
        grads = tr.gradients()
        errs = {}
        for k, g in refg.items():
            d = np.asarray(grads[k], np.float64).reshape(np.asarray(g).shape) - g
            errs[k] = (np.abs(d).max() / max(np.abs(g).max(), floor), np.linalg.norm(d) / max(np.linalg.norm(g), 1e-30))
        du1 = tr.block_gradient(0).cpu().numpy()
        u1 = tr.activation(0).cpu().numpy()
        du1_ref = np.concatenate([ref["du"][0][0], ref["du"][1][0]], axis=0) * (u1 > 0)
        gabs = tr.gabs.view(torch.float32).cpu().numpy()
        print(f"fwd {fp} bwd {bp}: loss rel err {abs(lv.item() - ref['loss']) / abs(ref['loss']):.2e}  "
              f"dU1 max err / max {np.abs(du1 - du1_ref).max() / np.abs(du1_ref).max():.2e}  "
              f"worst tensor max-rel {max(v[0] for v in errs.values()):.2e}  worst norm-rel "
              f"{max(v[1] for k, v in errs.items() if np.linalg.norm(refg[k]) > floor):.2e}  absmax per block {gabs}")
        print("   " + "  ".join(f"{k}:{v[0]:.1e}/{v[1]:.1e}" for k, v in errs.items()))


if __name__ == "__main__":
    main()
