"""Extracts the weights of the reference's shipped Keras checkpoint
(models/n_seconds/siamese__nseconds_3.0__filters_32__embed_64__drop_0.05__r_0.hdf5: real trained weights, older
architecture -- SURVEY.md F9) with this repo's pure-Python HDF5 reader and stores them, together with the CPU
oracle's embeddings / siamese outputs for a seeded input, as tests/golden/checkpoint_f32.npz.  The weights are a
realistic-distribution numerical fixture (moving variances 3e-7 .. 30, trained gammas); they are evaluated in the
CURRENT architecture (first pool 4), so outputs are not those of the network the checkpoint was trained as.
Run in the build container (needs /root/reference):  python tests/golden/make_checkpoint_fixture.py
"""
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import voicemap_oracle as O  # noqa: E402
from voicemap_b200.models import load_model  # noqa: E402

SRC = "/root/reference/models/n_seconds/siamese__nseconds_3.0__filters_32__embed_64__drop_0.05__r_0.hdf5"


def main():
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = load_model(SRC)
    w = {k: v.astype(np.float32) for k, v in m.encoder.weights.items()}
    # known-answer statistics recorded during the survey pin the reader itself
    assert abs(float(w["conv1_kernel"].mean()) - 0.000365) < 1e-6 and abs(float(w["conv1_kernel"].std()) - 0.115049) < 1e-6
    assert abs(float(m.head_weights["head_bias"][0]) + 2.692142) < 1e-6
    x = O.synthetic_clips(6, 12000, seed=2024, padded=True)
    e32 = O.encoder_forward(x, w, torch.float32)
    e64 = O.encoder_forward(x, w, torch.float64)
    prob, _ = O.siamese_head(e64[:3], e64[3:], m.head_weights["head_kernel"].reshape(-1).astype(np.float64),
                             float(m.head_weights["head_bias"][0]), "weighted_l1")
    out = {f"w_{k}": v for k, v in w.items()}
    out.update(head_kernel=m.head_weights["head_kernel"], head_bias=m.head_weights["head_bias"], x=x.astype(np.float32),
               emb32=e32, emb64=e64, prob64=prob)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "checkpoint_f32.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
