#!/bin/bash
# Short gpurun call while iterating on one kernel: block-level parity tests + timing breakdown (both precisions).
mkdir -p gpurun_out
LOG=gpurun_out/quick.log
: > $LOG
run() { echo "=== $*" >> $LOG; timeout ${TMO:-600} "$@" >> $LOG 2>&1; echo "--- exit $?" >> $LOG; }
run python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "${1:-block or golden or encoder_parity}"
run python tools/bringup.py --case time --n 256 --l 12000
run python tools/bringup.py --case time --n 256 --l 12000 --precision 1
tail -n 40 $LOG
