#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/time.log
: > $LOG
run() { echo "=== $*" >> $LOG; timeout ${TMO:-300} "$@" >> $LOG 2>&1; echo "--- exit $?" >> $LOG; }
run python tools/bringup.py --case time --n 256 --l 12000 --precision 2 --iters 20
run env VM_CONV3_NO_STORE=1 python tools/bringup.py --case time --n 256 --l 12000 --precision 2 --iters 20
run env VM_CONV3_NO_STORE=1 python tools/bringup.py --case time --n 256 --l 12000 --precision 3 --iters 20
grep -E "===|encoder|block|exit [1-9]" $LOG | cut -c1-200
