#!/usr/bin/env python
"""Per-block check of what the train-mode forward keeps for the backward pass (encoded activations, arg-max flags, window
extremes) against the fp64 oracle, at a given shape."""
import argparse, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import voicemap_oracle as O  # noqa: E402
from voicemap_b200.keras_compat import Adam  # noqa: E402
from voicemap_b200.models import build_siamese_net, get_baseline_convolutional_encoder  # noqa: E402
from voicemap_b200.training import TrainEngine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=2)
ap.add_argument("--length", type=int, default=12000)
ap.add_argument("--filters", type=int, default=128)
args = ap.parse_args()
n, length, f = args.pairs, args.length, args.filters
params = O.init_encoder_params(f, 64, seed=7, randomize_bn=False, random_bias=True)
rng = np.random.default_rng(8)
for i in range(1, 5):
    params[f"bn{i}_gamma"] = rng.uniform(-1.2, 1.5, params[f"bn{i}_gamma"].shape).astype(np.float32)
    params[f"bn{i}_beta"] = rng.normal(0, 0.2, params[f"bn{i}_beta"].shape).astype(np.float32)
x1, x2 = O.synthetic_clips(n, length, seed=111), O.synthetic_clips(n, length, seed=112)
y = (np.arange(n) >= n // 2).astype(np.float32)
enc = get_baseline_convolutional_encoder(f, 64, dropout=0.0)
enc.set_named_weights(params)
sia = build_siamese_net(enc, (length, 1))
sia.head_weights["head_kernel"][:] = 0.05
sia.head_weights["head_bias"][:] = -0.3
opt = Adam(clipnorm=1.0)
sia.compile(loss="contrastive_loss", optimizer=opt)
tr = TrainEngine(sia, opt, sia.loss, precision=3, bwd_precision=3)
hw, hb = sia.head_weights["head_kernel"].reshape(-1).copy(), sia.head_weights["head_bias"].copy()
lv, _ = tr.siamese_step(x1, x2, y, apply=False)
torch.cuda.synchronize()
masks = [[tr.relu_pattern(b)[br * n:(br + 1) * n].cpu().numpy().astype(np.float64) for b in range(4)] for br in range(2)]
ref = O.siamese_train_step_grads(params, hw, hb, x1, x2, y, loss="contrastive_loss", relu_masks=masks, pool_selects=[[tr.argmax_flags(b)[br * n:(br + 1) * n].cpu().numpy() for b in range(4)] for br in range(2)], gmax_selects=[tr.jstar[br * n:(br + 1) * n].cpu().numpy() for br in range(2)])
pools = (4, 2, 2, 2)
for b in range(4):
    u = tr.activation(b).cpu().numpy()
    flags = tr.argmax_flags(b).cpu().numpy()
    ext = tr.EXT[b].cpu().numpy()
    u_ref = np.concatenate([ref["u"][0][b], ref["u"][1][b]], axis=0)
    p = pools[b]
    lout = u_ref.shape[1] // p
    win = u_ref[:, :lout * p].reshape(u_ref.shape[0], lout, p, -1)
    neg = params[f"bn{b + 1}_gamma"] < 0
    pick = np.where(neg[None, None, :], win.argmin(axis=2), win.argmax(axis=2))
    want = np.zeros_like(win, dtype=bool)
    np.put_along_axis(want, pick[:, :, None, :], True, axis=2)
    want_full = np.zeros(u_ref.shape, dtype=bool)
    want_full[:, :lout * p] = want.reshape(u_ref.shape[0], lout * p, -1)
    ext_ref = np.where(neg[None, None, :], win.min(axis=2), win.max(axis=2))
    bad_flags = np.argwhere(flags != want_full)
    # ties (two equal values in a window) may legitimately pick differently only if values are equal: count non-tie errors
    print(f"block {b + 1}: u rel err {np.abs(u - u_ref).max() / np.abs(u_ref).max():.2e}  ext rel err "
          f"{np.abs(ext - ext_ref).max() / np.abs(ext_ref).max():.2e}  flag mismatches {len(bad_flags)} of {flags.size} "
          f"(flags set {int(flags.sum())}, expected {int(want_full.sum())})")
    if len(bad_flags):
        ls_ = np.unique(bad_flags[:, 1])
        print("   clips", np.unique(bad_flags[:, 0])[:8], "positions (first 24)", ls_[:24], "n positions", len(ls_), "pos mod 256 (first 24)", np.unique(ls_ % 256)[:24],
              "channels (first 16)", np.unique(bad_flags[:, 2])[:16], "n channels", len(np.unique(bad_flags[:, 2])))
refg = dict(ref["grads"], head_kernel=ref["head_w_grad"], head_bias=ref["head_b_grad"])
floor = 1e-3 * max(np.abs(np.asarray(g)).max() for g in refg.values())
grads = tr.gradients()
print("grad errs: " + "  ".join(f"{k}:{np.abs(np.asarray(grads[k], np.float64).reshape(np.asarray(g).shape) - g).max() / max(np.abs(g).max(), floor):.1e}" for k, g in refg.items()))
