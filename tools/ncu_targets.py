#!/usr/bin/env python
"""The launches that the committed ncu captures profile: ONE eval-mode forward of BASELINE config[1] (256 clips x 12000,
precision 2: conv1, conv3 x 3, gmax_dense) followed by ONE siamese training step of 64 pairs (both after an untimed
warm-up that ncu skips with --launch-skip).  Prints how many kernel launches the warm-up made.

    ncu --set full --clock-control none --import-source on --launch-skip <N> -o gpurun_out/r02_full \\
        python tools/ncu_targets.py            (see tools/gpu_ncu.sh)
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import voicemap_oracle as O  # noqa: E402  (seeded weights / inputs only)
from voicemap_b200.keras_compat import Adam  # noqa: E402
from voicemap_b200.models import build_siamese_net, get_baseline_convolutional_encoder  # noqa: E402
from voicemap_b200.training import TrainEngine  # noqa: E402
from voicemap_b200.utils import contrastive_loss  # noqa: E402

bwd = int(os.environ.get("VM_BWD", "1"))
params = O.init_encoder_params(128, 64, seed=0, randomize_bn=True, random_bias=True)
enc = get_baseline_convolutional_encoder(128, 64, dropout=0.0)
enc.set_named_weights(params)
enc.precision = 2
eng = enc._get_engine()
x = (O.WHITEN_RMS * torch.randn(256, 12000, generator=torch.Generator().manual_seed(1))).cuda()
out = torch.empty((256, 64), device="cuda")
enc2 = get_baseline_convolutional_encoder(128, 64, dropout=0.0)
sia = build_siamese_net(enc2, (12000, 1))
opt = Adam(clipnorm=1.0)
sia.compile(loss=contrastive_loss, optimizer=opt)
tr = TrainEngine(sia, opt, sia.loss, precision=3, bwd_precision=bwd)
x1, x2 = x[:64].contiguous(), x[64:128].contiguous()
y = torch.from_numpy((np.arange(64) >= 32).astype(np.float32)).cuda()
# warm-up (skipped by ncu): everything is allocated and packed afterwards
for _ in range(2):
    eng.forward(x, out=out)
    tr.siamese_step(x1, x2, y)
torch.cuda.synchronize()
print("warm-up done", flush=True)
eng.forward(x, out=out)
tr.siamese_step(x1, x2, y)
torch.cuda.synchronize()
