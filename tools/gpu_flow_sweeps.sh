#!/bin/bash
bash tools/gpu_flow_check.sh > /dev/null 2>&1
python tools/sweeps.py > gpurun_out/sweeps.log 2>&1
grep -E "=== |--- exit|val_|Epoch|reloaded" gpurun_out/flow.log | cut -c1-200 | tail -30
cut -c1-330 gpurun_out/sweeps.log | tail -14
