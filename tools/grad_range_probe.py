#!/usr/bin/env python
"""Dynamic range of the block-1 gradient dU1 at a given step shape, relative to the per-block scale (largest |s*dy|):
how much of it falls below fp16's normal / subnormal range once scaled.  Also prints per-tensor gradient errors of
the backward modes against the fp64 oracle at that shape."""
import argparse, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import voicemap_oracle as O  # noqa: E402
from voicemap_b200.keras_compat import Adam  # noqa: E402
from voicemap_b200.models import build_siamese_net, get_baseline_convolutional_encoder  # noqa: E402
from voicemap_b200.training import TrainEngine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=16)
ap.add_argument("--length", type=int, default=12000)
ap.add_argument("--modes", default="3,1")
ap.add_argument("--oracle", type=int, default=1)
args = ap.parse_args()
n, length = args.pairs, args.length
params = O.init_encoder_params(128, 64, seed=7, randomize_bn=False, random_bias=True)
rng = np.random.default_rng(8)
for i in range(1, 5):
    params[f"bn{i}_gamma"] = rng.uniform(-1.2, 1.5, params[f"bn{i}_gamma"].shape).astype(np.float32)
    params[f"bn{i}_beta"] = rng.normal(0, 0.2, params[f"bn{i}_beta"].shape).astype(np.float32)
x1, x2 = O.synthetic_clips(n, length, seed=111), O.synthetic_clips(n, length, seed=112)
y = (np.arange(n) >= n // 2).astype(np.float32)
ref = None
for bwd in [int(v) for v in args.modes.split(",")]:
    enc = get_baseline_convolutional_encoder(128, 64, dropout=0.0)
    enc.set_named_weights(params)
    sia = build_siamese_net(enc, (length, 1))
    sia.head_weights["head_kernel"][:] = 0.05
    sia.head_weights["head_bias"][:] = -0.3
    opt = Adam(clipnorm=1.0)
    sia.compile(loss="contrastive_loss", optimizer=opt)
    tr = TrainEngine(sia, opt, sia.loss, precision=3, bwd_precision=bwd)
    hw, hb = sia.head_weights["head_kernel"].reshape(-1).copy(), sia.head_weights["head_bias"].copy()
    lv, _ = tr.siamese_step(x1, x2, y, apply=False)
    torch.cuda.synchronize()
    gabs = tr.gabs.view(torch.float32).cpu().numpy()
    n_u = tr.U16[0].numel()
    scaled = tr.dU[:, :n_u].to(torch.float32).sum(dim=0).abs()      # as stored (scaled by the block's power of two)
    tot = scaled.sum().item()
    qs = torch.quantile(scaled[::97].contiguous(), torch.tensor([0.01, 0.1, 0.5, 0.9, 0.99, 0.9999], device=scaled.device)).cpu().numpy()
    sub = scaled < 6.1e-5
    print(f"bwd {bwd}: absmax/block {gabs}; scaled |dU1|: max {scaled.max().item():.3g} quantiles(1,10,50,90,99,99.99%) {qs}; "
          f"fraction of elements below fp16 normal range {sub.float().mean().item():.4f} carrying "
          f"{scaled[sub].sum().item() / tot:.4f} of the L1 mass; exactly zero {float((scaled == 0).float().mean()):.4f} "
          f"(relu-off fraction {1 - float(tr.relu_pattern(0).float().mean()):.4f})")
    if args.oracle:
        if ref is None:
            masks = [[tr.relu_pattern(b)[br * n:(br + 1) * n].cpu().numpy().astype(np.float64) for b in range(4)] for br in range(2)]
            ref = O.siamese_train_step_grads(params, hw, hb, x1, x2, y, loss="contrastive_loss", relu_masks=masks, pool_selects=[[tr.argmax_flags(b)[br * n:(br + 1) * n].cpu().numpy() for b in range(4)] for br in range(2)], gmax_selects=[tr.jstar[br * n:(br + 1) * n].cpu().numpy() for br in range(2)])
            del masks
        refg = dict(ref["grads"], head_kernel=ref["head_w_grad"], head_bias=ref["head_b_grad"])
        floor = 1e-3 * max(np.abs(np.asarray(g)).max() for g in refg.values())
        grads = tr.gradients()
        print("   " + "  ".join(f"{k}:{np.abs(np.asarray(grads[k], np.float64).reshape(np.asarray(g).shape) - g).max() / max(np.abs(g).max(), floor):.1e}" for k, g in refg.items()))
        print("   max|g| per tensor: " + "  ".join(f"{k}:{np.abs(g).max():.1e}" for k, g in refg.items()))
    del tr, sia, enc
    torch.cuda.empty_cache()
