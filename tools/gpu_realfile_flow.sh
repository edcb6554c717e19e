#!/bin/bash
# The training example on genuine FLAC files (decoder -> batcher -> preprocessing -> CUDA training step), first with the
# single background thread, then with forked batch producers.  Not yet run on a GPU box (written after round 1's GPU
# budget was spent): first thing to run in round 2.
mkdir -p gpurun_out
LOG=gpurun_out/realfile_flow.log
: > $LOG
run() { echo "=== $*" >> $LOG; timeout ${TMO:-300} "$@" >> $LOG 2>&1; echo "--- exit $?" >> $LOG; }
CORPUS=$(mktemp -d)
run python tools/make_flac_corpus.py $CORPUS --train-speakers 60 --dev-speakers 40 --files 4
export VOICEMAP_PATH=$CORPUS
OUT=$(mktemp -d); mkdir -p $OUT/logs $OUT/models
run python examples/train_siamese.py --epochs 2 --steps 40 --eval-tasks 40 --batchsize 64 --out $OUT
run python examples/train_siamese.py --epochs 2 --steps 40 --eval-tasks 40 --batchsize 64 --out $OUT --workers 8
run python examples/train_classifier.py --epochs 1 --steps 40 --eval-tasks 20 --out $OUT --workers 4
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then   # data parallel through the same script (gpurun --gpus 2)
  run python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
      examples/train_siamese.py --epochs 2 --steps 40 --eval-tasks 40 --batchsize 64 --out $OUT --workers 4
fi
tail -n 60 $LOG | cut -c1-300
