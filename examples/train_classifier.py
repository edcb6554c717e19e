#!/usr/bin/env python
"""Speaker-classification training through the reference's API (cf. experiments/train_classifier.py): the encoder
gets a Dense(num_speakers, softmax) head and is evaluated as an embedding on k-way n-shot tasks every epoch.

    python examples/train_classifier.py --synthetic --epochs 2 --steps 20
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from config import LIBRISPEECH_SAMPLING_RATE, PATH  # noqa: E402
from voicemap.librispeech import LibriSpeechDataset  # noqa: E402
from voicemap.models import get_baseline_convolutional_encoder  # noqa: E402
from voicemap.utils import BatchPreProcessor, NShotEvaluationCallback, preprocess_instances  # noqa: E402
from voicemap_b200.keras_compat import Adam, CSVLogger, Dense, Sequence, to_categorical  # noqa: E402


class ShuffledBatches(Sequence):
    """Batches of (clips, one-hot speaker) drawn from a LibriSpeechDataset, reshuffled every epoch."""

    def __init__(self, dataset, preprocessor, batchsize, seed=0):
        self.dataset, self.pre, self.batchsize = dataset, preprocessor, batchsize
        self.rng = np.random.default_rng(seed)
        self.order = self.rng.permutation(len(dataset))

    def __len__(self):
        return len(self.dataset) // self.batchsize

    def __getitem__(self, i):
        ids = self.order[i * self.batchsize:(i + 1) * self.batchsize]
        clips, labels = zip(*(self.dataset[int(j)] for j in ids))
        return self.pre((np.stack(clips)[:, :, np.newaxis], np.array(labels)[:, np.newaxis]))

    def on_epoch_end(self):
        self.order = self.rng.permutation(len(self.dataset))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--synthetic", action="store_true")
    ap.add_argument("--seconds", type=float, default=3)
    ap.add_argument("--downsampling", type=int, default=4)
    ap.add_argument("--batchsize", type=int, default=64)
    ap.add_argument("--filters", type=int, default=128)
    ap.add_argument("--embedding", type=int, default=64)
    ap.add_argument("--workers", type=int, default=1,
                    help="batch producer processes (the reference passes multiprocessing.cpu_count(), "
                         "experiments/train_siamese.py:71); 1 = one background thread")
    ap.add_argument("--epochs", type=int, default=50)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--eval-tasks", type=int, default=500)
    ap.add_argument("--out", default=PATH)
    args = ap.parse_args()

    if args.synthetic:
        from synthetic_speakers import SyntheticCorpus
        tr, va = SyntheticCorpus(40, 8, seed=0), SyntheticCorpus(12, 6, subset="synthetic-dev", seed=1)
        train = LibriSpeechDataset("synthetic", args.seconds, index=tr.index, reader=tr.reader)
        valid = LibriSpeechDataset("synthetic-dev", args.seconds, stochastic=False, index=va.index, reader=va.reader)
    else:
        train = LibriSpeechDataset(["train-clean-100", "train-clean-360"], args.seconds)
        valid = LibriSpeechDataset("dev-clean", args.seconds, stochastic=False)

    speakers = sorted(train.df["speaker_id"].unique())
    to_index = {s: i for i, s in enumerate(speakers)}
    pre = BatchPreProcessor("classifier", preprocess_instances(args.downsampling),
                            lambda y: to_categorical(np.array([to_index[s] for s in y[:, 0]]), len(speakers)))
    input_length = int(LIBRISPEECH_SAMPLING_RATE * args.seconds / args.downsampling)

    classifier = get_baseline_convolutional_encoder(args.filters, args.embedding, (input_length, 1))
    classifier.add(Dense(train.num_classes(), activation="softmax"))
    classifier.compile(loss="categorical_crossentropy", optimizer=Adam(clipnorm=1.), metrics=["accuracy"])
    classifier.summary()
    os.makedirs(os.path.join(args.out, "logs"), exist_ok=True)
    batches = ShuffledBatches(train, pre, args.batchsize)
    classifier.fit_generator(
        batches, steps_per_epoch=min(args.steps, len(batches)), epochs=args.epochs,
        workers=args.workers, use_multiprocessing=args.workers > 1,
        callbacks=[NShotEvaluationCallback(args.eval_tasks, 1, 5, valid, preprocessor=pre, mode="classifier"),
                   CSVLogger(os.path.join(args.out, "logs", "classifier.csv"))])


if __name__ == "__main__":
    main()
