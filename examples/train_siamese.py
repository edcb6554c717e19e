#!/usr/bin/env python
"""Siamese verification training through the reference's API (cf. experiments/train_siamese.py of the reference):
LibriSpeechDataset verification batches -> BatchPreProcessor(downsample x4 + whiten) -> shared 1D-CNN encoder ->
uniform_euclidean head -> Adam(clipnorm=1) with the n-shot callback, CSV log, checkpoint and LR schedule.

    python examples/train_siamese.py --synthetic --epochs 2 --steps 20
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from config import LIBRISPEECH_SAMPLING_RATE, PATH  # noqa: E402
from voicemap.librispeech import LibriSpeechDataset  # noqa: E402
from voicemap.models import build_siamese_net, get_baseline_convolutional_encoder  # noqa: E402
from voicemap.utils import BatchPreProcessor, NShotEvaluationCallback, contrastive_loss, preprocess_instances  # noqa: E402
from voicemap_b200.keras_compat import Adam, CSVLogger, ModelCheckpoint, ReduceLROnPlateau  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--synthetic", action="store_true", help="use the synthetic corpus instead of LibriSpeech on disk")
    ap.add_argument("--seconds", type=float, default=3)
    ap.add_argument("--downsampling", type=int, default=4)
    ap.add_argument("--batchsize", type=int, default=64)
    ap.add_argument("--filters", type=int, default=128)
    ap.add_argument("--embedding", type=int, default=64)
    ap.add_argument("--dropout", type=float, default=0.0)
    ap.add_argument("--workers", type=int, default=1,
                    help="batch producer processes (the reference passes multiprocessing.cpu_count(), "
                         "experiments/train_siamese.py:71); 1 = one background thread")
    ap.add_argument("--epochs", type=int, default=50)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--eval-tasks", type=int, default=500)
    ap.add_argument("--k-way", type=int, default=5)
    ap.add_argument("--loss", default="binary_crossentropy", choices=["binary_crossentropy", "contrastive_loss"])
    ap.add_argument("--out", default=PATH)
    args = ap.parse_args()

    # data parallel: `python -m torch.distributed.run --nproc-per-node N examples/train_siamese.py ...` gives every rank
    # 1/N of each batch; fit_generator then shares BatchNorm statistics and gradients and starts from rank 0's weights
    rank, world = 0, 1
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch
        from voicemap_b200 import parallel
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        rank, world = parallel.init_from_env("nccl", torch.device("cuda", local))
        args.batchsize = max(2, args.batchsize // world // 2 * 2)

    if args.synthetic:
        from synthetic_speakers import SyntheticCorpus
        # the reference's differing-pair draw excludes every speaker of its first draw: keep speakers >> batchsize / 2
        tr, va = SyntheticCorpus(120, 4, seed=0), SyntheticCorpus(80, 3, subset="synthetic-dev", seed=1)
        train = LibriSpeechDataset("synthetic", args.seconds, pad=True, index=tr.index, reader=tr.reader)
        valid = LibriSpeechDataset("synthetic-dev", args.seconds, stochastic=False, pad=True, index=va.index,
                                   reader=va.reader)
    else:
        train = LibriSpeechDataset(["train-clean-100", "train-clean-360"], args.seconds, pad=True)
        valid = LibriSpeechDataset("dev-clean", args.seconds, stochastic=False, pad=True)

    input_length = int(LIBRISPEECH_SAMPLING_RATE * args.seconds / args.downsampling)
    tag = "siamese__filters_{}__embed_{}__drop_{}".format(args.filters, args.embedding, args.dropout)
    pre = BatchPreProcessor("siamese", preprocess_instances(args.downsampling))
    train_batches = (pre(b) for b in train.yield_verification_batches(args.batchsize))
    valid_batches = (pre(b) for b in valid.yield_verification_batches(args.batchsize))

    encoder = get_baseline_convolutional_encoder(args.filters, args.embedding, dropout=args.dropout)
    siamese = build_siamese_net(encoder, (input_length, 1), distance_metric="uniform_euclidean")
    siamese.compile(loss=contrastive_loss if args.loss == "contrastive_loss" else args.loss,
                    optimizer=Adam(clipnorm=1.), metrics=["accuracy"])
    siamese.summary()

    for d in ("logs", "models"):
        os.makedirs(os.path.join(args.out, d), exist_ok=True)
    monitor = "val_1-shot_acc"
    siamese.fit_generator(
        train_batches, steps_per_epoch=args.steps, epochs=args.epochs,
        validation_data=valid_batches, validation_steps=max(1, args.steps // 5),
        workers=args.workers, use_multiprocessing=args.workers > 1,
        verbose=1 if rank == 0 else 0,
        callbacks=[NShotEvaluationCallback(max(1, args.eval_tasks // world), 1, args.k_way, valid, preprocessor=pre)] + ([
            # files are written by rank 0 only; the metric they follow is pooled over ranks by the callback above
            CSVLogger(os.path.join(args.out, "logs", tag + ".csv")),
            ModelCheckpoint(os.path.join(args.out, "models", tag + ".hdf5"), monitor=monitor, mode="max",
                            save_best_only=True, verbose=True),
        ] if rank == 0 else []) + [ReduceLROnPlateau(monitor=monitor, mode="max", verbose=1 if rank == 0 else 0)])


if __name__ == "__main__":
    main()
