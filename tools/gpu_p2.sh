#!/bin/bash
# precision 2 check: GPU suite, smoke, bench (both precisions), ncu launch list + full capture of the conv kernels
mkdir -p gpurun_out
LOG=gpurun_out/p2.log
: > $LOG
run() { echo "=== $*" >> $LOG; timeout ${TMO:-600} "$@" >> $LOG 2>&1; echo "--- exit $?" >> $LOG; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $LOG 2>&1
run python -m pytest tests -x -q -m gpu
run python -c "import __graft_entry__ as g; g.smoke()"
run python bench.py --steps 100 --warmup 5
run python bench.py --steps 100 --warmup 5 --precision 3 --no-cpu-baseline
if [ "$1" == "ncu" ]; then
  run ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_p2.csv \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline
  TMO=900 run ncu --set full --clock-control none --import-source on -k regex:conv -s 8 -c 4 -f -o gpurun_out/prof_p2 \
      python tools/bringup.py --case time --n 256 --l 12000 --iters 1 --precision 2
fi
tail -n 60 $LOG | cut -c1-600
