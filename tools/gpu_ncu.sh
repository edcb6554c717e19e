#!/bin/bash
# ncu evidence of a round: (1) launch list (durations) of one eval forward + one training step, (2) --set full capture of
# the same launches, exported as text sections (gpurun_out/<tag>_ncu_details.txt) next to the .ncu-rep.
#   tools/gpu_ncu.sh r02            (on the GPU box, through gpurun)
TAG=${1:-r02}
mkdir -p gpurun_out
# count the warm-up launches: a first pass with the cheap metric lists everything
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_ncu_all_launches.csv \
    python tools/ncu_targets.py > gpurun_out/${TAG}_ncu_run.log 2>&1
TOTAL=$(grep -c '^"[0-9]' gpurun_out/${TAG}_ncu_all_launches.csv)
# the profiled tail = last forward (5 kernels + pack if weights changed) + last training step: find its first launch = the
# second-to-last conv1 eval launch
python - "$TAG" <<'PY'
import csv, sys
tag = sys.argv[1]
rows = [r for r in csv.reader(l for l in open(f"gpurun_out/{tag}_ncu_all_launches.csv") if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
name = hdr.index("Kernel Name")
idx = [i for i, r in enumerate(rows) if "conv1_kernel" in r[name]]
# launches of conv1: [fwd, train] per iteration; the tail starts at the last eval forward's conv1
start = idx[-2]
import re
pat = re.compile("conv1_kernel|conv3_kernel|wgrad|bn_relu_bwd|bn_bwd_reduce|bn_pool_fwd|gmax_dense|siamese_head|bn_gmax_fwd")
# --launch-skip counts only launches that pass the -k filter of the full capture
open(f"gpurun_out/{tag}_ncu_skip.txt", "w").write(str(sum(1 for r in rows[:start] if pat.search(r[name]))))
val = hdr.index("Metric Value")
with open(f"gpurun_out/{tag}_ncu_tail_launches.csv", "w") as f:
    f.write("index,kernel,grid,block,duration_us\n")
    for i, r in enumerate(rows[start:]):
        f.write(f'{i},"{r[name].split("(")[0]}","{r[hdr.index("Grid Size")]}","{r[hdr.index("Block Size")]}",{float(r[val]) / 1e3:.2f}\n')
print("tail starts at launch", start, "of", len(rows))
PY
SKIP=$(cat gpurun_out/${TAG}_ncu_skip.txt)
# the report itself (~80 MB) stays on the box: gpurun brings back at most 64 MiB, so only the text / csv exports travel
REP=/tmp/${TAG}_full
ncu --set full --clock-control none --import-source on --launch-skip $SKIP \
    -k regex:'conv1_kernel|conv3_kernel|wgrad|bn_relu_bwd|bn_bwd_reduce|bn_pool_fwd|gmax_dense|siamese_head|bn_gmax_fwd' -f -o $REP \
    python tools/ncu_targets.py >> gpurun_out/${TAG}_ncu_run.log 2>&1
ncu -i $REP.ncu-rep --page details > gpurun_out/${TAG}_ncu_details.txt 2>&1
ncu -i $REP.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_raw.csv 2>&1
ls -la gpurun_out/${TAG}_* | tail -8
tail -3 gpurun_out/${TAG}_ncu_run.log
