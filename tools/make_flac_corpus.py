"""Writes the synthetic speakers of examples/synthetic_speakers.py as a LibriSpeech-shaped tree of genuine FLAC files
(<root>/data/LibriSpeech/<subset>/<speaker>/<chapter>/<utterance>.flac + SPEAKERS.TXT), so that the examples and the
batcher run on the real file path -- C decoder, header indexing, fragment reads, producer processes -- without
LibriSpeech itself, which this image does not have.  16-bit, 16 kHz, VERBATIM subframes (the test-side writer in
tests/flac_writer.py; fast, uncompressed).

    python tools/make_flac_corpus.py /tmp/corpus            # train-clean-100, train-clean-360, dev-clean
    VOICEMAP_PATH=/tmp/corpus python examples/train_siamese.py --workers 8 --epochs 2 --steps 50
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "examples")]
from flac_writer import encode_flac_quick  # noqa: E402
from synthetic_speakers import SyntheticCorpus  # noqa: E402


def write_subset(root, subset, n_speakers, files_per_speaker, seed, first_id, lines):
    corpus = SyntheticCorpus(n_speakers, files_per_speaker, subset=subset, seed=seed)
    written = 0
    for row in corpus.index.itertuples():
        speaker = first_id + (row.id - 1000)
        utterance = int(row.filepath.split("/")[3])
        folder = os.path.join(root, "data", "LibriSpeech", subset, str(speaker), "1")
        os.makedirs(folder, exist_ok=True)
        samples, _ = corpus.reader(row.filepath)
        pcm = np.clip(np.round(samples * 32768.0), -32768, 32767).astype(np.int64)
        with open(os.path.join(folder, "{}-1-{:04d}.flac".format(speaker, utterance)), "wb") as handle:
            handle.write(encode_flac_quick(pcm))
        written += 1
    for s in range(n_speakers):
        sex = corpus.index[corpus.index["id"] == 1000 + s]["sex"].iloc[0]
        lines.append("{:<5}| {} | {:<16} | 25.00 | Synthetic speaker {}".format(first_id + s, sex, subset, s))
    return written


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("root")
    ap.add_argument("--train-speakers", type=int, default=120, help="per training subset")
    ap.add_argument("--dev-speakers", type=int, default=60)
    ap.add_argument("--files", type=int, default=4, help="utterances per speaker (3.5 - 5 s each)")
    args = ap.parse_args()
    lines = ["; synthetic speakers, see examples/synthetic_speakers.py", ";ID  |SEX| SUBSET           |MINUTES| NAME"]
    total = 0
    total += write_subset(args.root, "train-clean-100", args.train_speakers, args.files, 0, 1000, lines)
    total += write_subset(args.root, "train-clean-360", args.train_speakers, args.files, 1, 3000, lines)
    total += write_subset(args.root, "dev-clean", args.dev_speakers, args.files, 2, 5000, lines)
    with open(os.path.join(args.root, "data", "LibriSpeech", "SPEAKERS.TXT"), "w") as handle:
        handle.write("\n".join(lines) + "\n")
    print("wrote {} utterances under {}".format(total, os.path.join(args.root, "data", "LibriSpeech")))


if __name__ == "__main__":
    main()
