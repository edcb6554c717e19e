#!/bin/bash
# User-flow check on a GPU box: the example training script end to end (verification batches -> fit_generator with the
# n-shot callback, CSV log, HDF5 checkpoint, LR schedule), reloading the checkpoint it wrote, and small-batch train timing.
mkdir -p gpurun_out
LOG=gpurun_out/flow.log
: > $LOG
run() { echo "=== $*" >> $LOG; timeout ${TMO:-600} "$@" >> $LOG 2>&1; echo "--- exit $?" >> $LOG; }
OUT=$(mktemp -d)
run python examples/train_siamese.py --synthetic --epochs 2 --steps 12 --eval-tasks 40 --batchsize 32 --filters 32 --out $OUT
run python - <<PY
import glob, numpy as np, sys
sys.path.insert(0, ".")
from voicemap_b200.models import load_model
path = glob.glob("$OUT/models/*.hdf5")[0]
m = load_model(path)
x = (0.038 * np.random.default_rng(0).normal(size=(4, 12000, 1))).astype(np.float32)
p = m.predict([x[:2], x[2:]])
print("reloaded", path.split("/")[-1], type(m).__name__, "first_pool", m.encoder.first_pool, "prob", p.reshape(-1))
print(open(glob.glob("$OUT/logs/*.csv")[0]).read())
PY
run python examples/train_classifier.py --synthetic --epochs 2 --steps 10 --filters 32 --out $OUT
run python tools/train_bench.py --pairs-per-gpu 16 --steps 20
run python tools/train_bench.py --pairs-per-gpu 32 --steps 20
tail -n 60 $LOG
