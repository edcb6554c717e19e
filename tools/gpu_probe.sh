#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/probe.log
: > $LOG
run() { echo "=== $*" >> $LOG; timeout ${TMO:-300} "$@" >> $LOG 2>&1; echo "--- exit $?" >> $LOG; }
run python tools/bringup.py --case time --n 256 --l 12000
VM_PROBE_F8=1 run env VM_PROBE_F8=1 python tools/bringup.py --case time --n 256 --l 12000
VM_PROBE_F8=2 run env VM_PROBE_F8=2 python tools/bringup.py --case time --n 256 --l 12000
run python tools/h2d_probe.py
tail -n 60 $LOG | cut -c1-300
