#!/bin/bash
# One gpurun call: whole GPU suite + smoke + timing + bench + train bench (+ optional sweeps)
mkdir -p gpurun_out
LOG=gpurun_out/full.log
: > $LOG
run() { echo "=== $*" >> $LOG; timeout ${TMO:-900} "$@" >> $LOG 2>&1; echo "--- exit $?" >> $LOG; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $LOG 2>&1
run python -m pytest tests -q -m gpu
run python -c "import __graft_entry__ as g; g.smoke()"
run python tools/bringup.py --case time --n 256 --l 12000
run python tools/bringup.py --case time --n 256 --l 12000 --precision 1
run python bench.py --warmup 5
run python tools/train_bench.py --pairs-per-gpu 64 --steps 10
run python tools/train_bench.py --pairs-per-gpu 64 --steps 10 --bwd-precision 1
if [ "$1" == "sweeps" ]; then run python tools/sweeps.py; fi
tail -n 60 $LOG
