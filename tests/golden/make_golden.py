"""Generates tests/golden/encoder_f128_e64.npz: seeded inputs/weights recipe + the CPU oracle's outputs (fp32 and
fp64) for the hot path.  The reference itself (Python 2.7 / Keras 2.2.2 / TF 1.10) cannot run in this image, so
these vectors pin the *oracle restatement* (oracle/voicemap_oracle.py), not Keras -- "parity unpinned", see
DESIGN.md.  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import voicemap_oracle as O  # noqa: E402


def main():
    filters, emb, n, length = 128, 64, 6, 3000
    params = O.init_encoder_params(filters, emb, seed=11, randomize_bn=True, random_bias=True)
    x = O.synthetic_clips(n, length, seed=4321, padded=True)
    e32, inter32, g32, _ = O.encoder_forward(x, params, torch.float32, return_intermediates=True)
    e64 = O.encoder_forward(x, params, torch.float64)
    # head scaled to the embedding distances so that the sigmoid is not saturated (a saturated p makes the
    # Keras 1e-7 clip, evaluated in fp32 on the GPU and in fp64 here, the only thing being compared)
    d0 = np.sqrt(np.sum(np.square(e64[:3] - e64[3:]), axis=-1))
    head_w, head_b = np.float32(2.0 / np.median(d0)), np.float32(-1.5)
    prob, dist = O.siamese_head(e64[:3], e64[3:], head_w, head_b)
    y = np.array([[0.0], [0.0], [1.0]])
    out = dict(filters=filters, emb=emb, n=n, length=length, param_seed=11, input_seed=4321,
               x=x.astype(np.float32), emb32=e32, emb64=e64, gmax32=g32,
               block1_sample=inter32[0][:, :40, :].copy(), block2_sample=inter32[1][:, :20, :].copy(),
               head_w=head_w, head_b=head_b, prob64=prob, dist64=dist, y=y,
               contrastive64=O.contrastive_loss(y, prob), bce64=O.binary_crossentropy(y, prob))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "encoder_f128_e64.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
