#!/bin/bash
# Short final check of a round: the newest GPU tests first, then the whole GPU suite, smoke and the bench line.
mkdir -p gpurun_out
LOG=gpurun_out/final.log
: > $LOG
run() { echo "=== $*" >> $LOG; timeout ${TMO:-240} "$@" >> $LOG 2>&1; echo "--- exit $?" >> $LOG; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $LOG 2>&1
run python -m pytest tests/test_reference_golden.py -q -m gpu
run python -m pytest tests -x -q -m gpu
run python -c "import __graft_entry__ as g; g.smoke()"
run python bench.py
tail -n 40 $LOG | cut -c1-600
