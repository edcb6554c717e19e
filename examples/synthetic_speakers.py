"""A LibriSpeech-shaped synthetic corpus (no audio ships with the reference or this image): every "speaker" is a
random mixture of a few harmonics with speaker-specific formant-like envelopes plus noise, so that verification
and n-shot tasks are learnable.  Produces the index DataFrame and reader that ``LibriSpeechDataset`` accepts."""
import numpy as np
import pandas as pd

RATE = 16000


class SyntheticCorpus:
    def __init__(self, n_speakers=40, files_per_speaker=6, seconds=(3.5, 5.0), subset="synthetic", seed=0):
        self.rng = np.random.default_rng(seed)
        self.voices = {}
        self._cache = {}
        rows = []
        for s in range(n_speakers):
            f0 = self.rng.uniform(90, 260)
            amps = self.rng.dirichlet(np.ones(6)) * self.rng.uniform(0.5, 1.5)
            self.voices[1000 + s] = (f0, amps, self.rng.uniform(0.1, 0.4))
            for j in range(files_per_speaker):
                length = int(RATE * self.rng.uniform(*seconds))
                rows.append(dict(id=1000 + s, sex="M" if f0 < 165 else "F", subset=subset, minutes=25.0,
                                 name=f"speaker{s}", filepath=f"synthetic://{1000 + s}/{j}/{length}", length=length,
                                 seconds=length / RATE))
        self.index = pd.DataFrame(rows)

    def reader(self, path):
        if path in self._cache:
            return self._cache[path], RATE
        audio, _ = self._synthesise(path)
        if len(self._cache) < 4096:          # a few hundred MB at most: evaluation loops re-read the same files
            audio.setflags(write=False)      # shared between readers from now on
            self._cache[path] = audio
        return audio, RATE

    def _synthesise(self, path):
        _, _, spk, j, length = path.split("/")
        spk, j, length = int(spk), int(j), int(length)
        rng = np.random.default_rng(spk * 1000 + j)
        f0, amps, noise = self.voices[spk]
        t = np.arange(length) / RATE
        vib = 1.0 + 0.01 * np.sin(2 * np.pi * rng.uniform(3, 7) * t)
        x = sum(a * np.sin(2 * np.pi * f0 * (k + 1) * vib * t + rng.uniform(0, 6.28)) for k, a in enumerate(amps))
        x = x * (0.6 + 0.4 * np.sin(2 * np.pi * rng.uniform(1, 4) * t)) + noise * rng.normal(size=length)
        return 0.05 * x, RATE
