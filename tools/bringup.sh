#!/bin/bash
# One gpurun call: isolated bring-up experiments, each under its own timeout.
mkdir -p gpurun_out
LOG=gpurun_out/bringup.log
: > $LOG
run() { echo "=== $*" >> $LOG; timeout 180 "$@" >> $LOG 2>&1; echo "--- exit $?" >> $LOG; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $LOG 2>&1
run python tools/bringup.py --case conv1 --n 2 --l 2048
run python tools/bringup.py --case conv1 --n 3 --l 1999
run python tools/bringup.py --case conv3 --block 2 --n 2 --l 2048
run python tools/bringup.py --case conv3 --block 3 --n 3 --l 1999
run python tools/bringup.py --case conv3 --block 4 --n 2 --l 2048
run python tools/bringup.py --case encoder --n 8 --l 12000
run python tools/bringup.py --case encoder --n 5 --l 11999
run python tools/bringup.py --case encoder --n 8 --l 12000 --precision 1
run python tools/bringup.py --case time --n 256 --l 12000
run python tools/bringup.py --case time --n 256 --l 12000 --precision 1
run python tools/bringup.py --case time --n 64 --l 48000
tail -n 80 $LOG
