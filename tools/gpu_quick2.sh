#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/quick2.log
: > $LOG
run() { echo "=== $*" >> $LOG; timeout ${TMO:-300} "$@" >> $LOG 2>&1; echo "--- exit $?" >> $LOG; }
run python -m pytest tests/test_gpu_parity.py -x -q -m gpu
run python tools/bringup.py --case time --n 256 --l 12000 --precision 2 --iters 20
run python tools/bringup.py --case time --n 256 --l 12000 --precision 3 --iters 20
tail -n 50 $LOG | cut -c1-300
