#!/bin/bash
# multi-GPU validation: bench.py under torchrun (weak scaling, no collective on the data path) and the DDP train step
mkdir -p gpurun_out
LOG=gpurun_out/multi.log
: > $LOG
N=${1:-2}
run() { echo "=== $*" >> $LOG; timeout ${TMO:-600} "$@" >> $LOG 2>&1; echo "--- exit $?" >> $LOG; }
nvidia-smi --query-gpu=index,name --format=csv >> $LOG 2>&1
run python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 5
run python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1
run python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/train_bench.py --steps 10
run python tools/train_bench.py --pairs-per-gpu $((128 / N)) --steps 10
run python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 tools/ddp_parity.py
tail -n 40 $LOG
