#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/train.log
: > $LOG
run() { echo "=== $*" >> $LOG; timeout ${TMO:-600} "$@" >> $LOG 2>&1; echo "--- exit $?" >> $LOG; }
run python -m pytest tests/test_gpu_train.py -q -m gpu -s
run python -m pytest tests/test_gpu_parity.py -q -m gpu
run python bench.py --steps 100 --warmup 5 --no-cpu-baseline
tail -n 150 $LOG
