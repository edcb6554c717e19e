#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/train.log
: > $LOG
run() { echo "=== $*" >> $LOG; timeout ${TMO:-600} "$@" >> $LOG 2>&1; echo "--- exit $?" >> $LOG; }
run python -m pytest tests/test_gpu_train.py -q -m gpu -s
run python tools/train_bench.py --pairs-per-gpu 64 --steps 10
run python tools/train_bench.py --pairs-per-gpu 16 --steps 10
tail -n 150 $LOG
