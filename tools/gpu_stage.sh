#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/stage.log
: > $LOG
run() { echo "=== $*" >> $LOG; timeout ${TMO:-240} "$@" >> $LOG 2>&1; echo "--- exit $?" >> $LOG; }
run python -m pytest tests -x -q -m gpu
run python tools/predict_numpy_bench.py
run python -c "import __graft_entry__ as g; g.smoke()"
run python bench.py --no-cpu-baseline
tail -n 40 $LOG | cut -c1-400
