#!/bin/bash
# One generic GPU-box job runner: tools/gpu_job.sh NAME 'cmd 1' 'cmd 2' ...
# Every command runs under its own timeout (TMO seconds, default 600) with output appended to gpurun_out/NAME.log;
# the log tail is printed at the end so that the gpurun call's own output shows the verdicts.
NAME=$1; shift
mkdir -p gpurun_out
LOG=gpurun_out/$NAME.log
: > "$LOG"
for cmd in "$@"; do
  echo "=== $cmd" >> "$LOG"
  timeout "${TMO:-600}" bash -c "$cmd" >> "$LOG" 2>&1
  echo "--- exit $?" >> "$LOG"
done
tail -n "${TAIL:-120}" "$LOG"
