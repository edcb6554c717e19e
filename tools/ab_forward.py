#!/usr/bin/env python
"""Same-box A/B of the eval forward (256 x 12000, precision 2): device-resident step time, alternating rounds, for the
variants selected by environment switches of the library (each variant = a fresh subprocess, since the switch is read
once).  Usage: python tools/ab_forward.py [ENVVAR | lib=PATH ...]   (baseline = no variable set; lib=PATH loads another
build of libvoicemap_b200.so, e.g. one built from an earlier commit, with the symbols it has)"""
import json
import os
import subprocess
import sys

CHILD = r'''
import os, sys, torch, numpy as np, ctypes
sys.path.insert(0, ".")
from voicemap_b200 import _lib
if os.environ.get("VOICEMAP_AB_LIB"):
    _lib.LIB_PATH = os.environ["VOICEMAP_AB_LIB"]
    probe = ctypes.CDLL(_lib.LIB_PATH)
    _lib.SIGNATURES = {k: v for k, v in _lib.SIGNATURES.items() if hasattr(probe, k)}
from oracle import voicemap_oracle as O
from voicemap_b200.engine import EncoderEngine
eng = EncoderEngine(128, 64, precision=2)
eng.set_weights(O.init_encoder_params(128, 64, seed=0, randomize_bn=True, random_bias=True))
g = torch.Generator().manual_seed(1)
sets = [(O.WHITEN_RMS * torch.randn(256, 12000, generator=g)).cuda() for _ in range(14)]
out = torch.empty((256, 64), device="cuda")
for i in range(10): eng.forward(sets[i % 14], out=out)
torch.cuda.synchronize()
best = []
for rep in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(40): eng.forward(sets[i % 14], out=out)
    b.record(); torch.cuda.synchronize()
    best.append(a.elapsed_time(b) / 40)
print(min(best), float(np.median(best)))
'''

def run(var):
    env = dict(os.environ)
    if var and var.startswith("lib="):
        env["VOICEMAP_AB_LIB"] = os.path.abspath(var[4:])
    elif var:
        env[var] = "1"
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    if r.returncode != 0:
        return r.stderr[-400:]
    return [float(v) for v in r.stdout.split()[-2:]]

variants = [None] + sys.argv[1:]
for rnd in range(3):
    for v in variants:
        print(json.dumps({"round": rnd, "variant": v or "default", "ms_best_median": run(v)}), flush=True)
