"""Seeded recipes shared by make_reference_golden.py (which feeds them to the reference's own code) and
tests/test_reference_golden.py (which feeds them to the oracle, the host utilities and the CUDA path)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from oracle import voicemap_oracle as O  # noqa: E402  (seeded generators; tests may import the oracle)

FIXTURE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_executed.npz")

# name -> (filters, embedding, clips, samples per clip, parameter seed, input seed)
ENCODER_CASES = {
    "f16": (16, 32, 6, 1999, 21, 31),        # ragged length: every pooling stage drops a tail
    "f128": (128, 64, 4, 3000, 22, 32),      # the benchmark's width
}


def encoder_inputs(name):
    filters, emb, n, length, pseed, xseed = ENCODER_CASES[name]
    params = O.init_encoder_params(filters, emb, seed=pseed, randomize_bn=True, random_bias=True)
    x = O.synthetic_clips(n, length, seed=xseed, padded=(name == "f16"))
    return params, np.asarray(x, dtype=np.float32)                 # (n, samples, 1)


def keras_weight_list(params):
    """The encoder's weights in Keras `get_weights()` order: per block kernel, bias, gamma, beta, moving mean, moving
    variance; then the Dense kernel and bias."""
    out = []
    for i in (1, 2, 3, 4):
        out += [params[f"conv{i}_kernel"], params[f"conv{i}_bias"], params[f"bn{i}_gamma"], params[f"bn{i}_beta"],
                params[f"bn{i}_mean"], params[f"bn{i}_var"]]
    return out + [params["dense_kernel"], params["dense_bias"]]


def head_weights(metric, emb, scale):
    """Dense(1) of the siamese head, scaled so that the sigmoid is not saturated."""
    rng = np.random.default_rng(77)
    if metric == "uniform_euclidean":
        return np.array([[2.0 / scale]]), np.array([-1.5])
    return rng.uniform(0.5, 1.5, size=(emb, 1)) * (4.0 / (scale * emb)), np.array([-1.0])


def classifier_head(emb, classes=7):
    rng = np.random.default_rng(78)
    return rng.normal(0, 0.5, size=(emb, classes)), rng.normal(0, 0.1, size=(classes,))


def raw_clips(n=3, samples=4001, seed=5):
    """Raw 16 kHz style input for preprocess_instances: (n, samples, 1) float64 with per-clip offsets and gains."""
    rng = np.random.default_rng(seed)
    return rng.normal(0, 1, size=(n, samples, 1)) * rng.uniform(0.01, 0.3, size=(n, 1, 1)) + rng.normal(0, 0.05, (n, 1, 1))


class TaskDataset:
    """Deterministic stand-in for LibriSpeechDataset.build_n_shot_task (voicemap/librispeech.py:204-240): `speakers`
    synthetic voices = fixed random waveforms; a clip is its speaker's waveform plus noise strong enough that tasks
    are not all solved.  The sequence of tasks depends only on `seed`."""

    def __init__(self, seed, samples=4096, speakers=12, noise=1.0):
        self.rng = np.random.default_rng(seed)
        self.voices = np.random.default_rng(1000 + seed).normal(0, 0.05, size=(speakers, samples))
        self.noise = noise * 0.05
        self.samples = samples

    def _clip(self, speaker):
        return self.voices[speaker] + self.rng.normal(0, self.noise, size=self.samples)

    def build_n_shot_task(self, k, n=1):
        speakers = self.rng.choice(len(self.voices), size=k, replace=False)
        query = (self._clip(speakers[0]), speakers[0])
        support = np.stack([self._clip(s) for s in speakers for _ in range(n)])
        labels = np.repeat(speakers, n)
        return query, (support, labels)


# (network_type, n, k, distance, tasks, dataset seed, noise level)
NSHOT_CASES = [
    ("siamese", 1, 5, "euclidean", 24, 1, 0.3),
    ("siamese", 3, 4, "euclidean", 16, 2, 0.5),
    ("siamese", 2, 5, "cosine", 16, 3, 0.5),
    ("siamese", 2, 3, "dot_product", 16, 4, 0.5),
    ("classifier", 2, 4, "euclidean", 12, 5, 0.5),
]
NSHOT_DOWNSAMPLING = 4
