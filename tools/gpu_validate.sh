#!/bin/bash
# One gpurun call: GPU suite, smoke, both bench arms, train-step timings.
mkdir -p gpurun_out
LOG=gpurun_out/validate.log
: > $LOG
run() { echo "=== $*" >> $LOG; timeout ${TMO:-600} "$@" >> $LOG 2>&1; echo "--- exit $?" >> $LOG; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $LOG 2>&1
run python -m pytest tests -x -q -m gpu
run python -c "import __graft_entry__ as g; g.smoke()"
run python bench.py
run python bench.py --impl reference --steps 2 --warmup 1
run python tools/train_bench.py --pairs-per-gpu 16 --steps 30
run python tools/train_bench.py --pairs-per-gpu 64 --steps 30
tail -n 80 $LOG | cut -c1-400
