#!/bin/bash
# One gpurun call: GPU test-suite, smoke, bench, timing breakdown, and (optionally) ncu captures.
# usage: bash tools/gpu_check.sh [ncu]
mkdir -p gpurun_out
LOG=gpurun_out/check.log
: > $LOG
run() { echo "=== $*" >> $LOG; timeout ${TMO:-600} "$@" >> $LOG 2>&1; echo "--- exit $?" >> $LOG; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $LOG 2>&1
run python -m pytest tests -x -q -m gpu
run python -c "import __graft_entry__ as g; g.smoke()"
run python tools/bringup.py --case time --n 256 --l 12000
run python tools/bringup.py --case time --n 256 --l 12000 --precision 1
run python bench.py --steps 30 --warmup 5
run python bench.py --impl reference --steps 3 --warmup 1
if [ "$1" == "ncu" ]; then
  TMO=900 run ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline
  TMO=900 run ncu --set full --clock-control none --import-source on -k regex:conv -s 8 -c 4 -f -o gpurun_out/prof \
      python tools/bringup.py --case time --n 256 --l 12000 --iters 1
fi
tail -n 120 $LOG
