"""The reference's OWN test file (/root/reference/tests/tests.py, read where it lies, unchanged) run against this
repo's drop-in packages: its imports -- `voicemap.utils.whiten`, `voicemap.librispeech.LibriSpeechDataset`,
`config.PATH` -- resolve to voicemap_b200.  It exercises the batcher on a LibriSpeech tree on disk, so a miniature
corpus of genuine FLAC files is written first (21 speakers x 6 utterances longer than 3 s, including the one file the
whitening test opens by name) and decoded by our own C decoder.

Two environment shims, no edits: `soundfile` (absent from the image) is served by voicemap_b200.audio_io, and
`pandas.value_counts` (removed in pandas 3, the reference pins 0.23) by its Series equivalent.

Runs where the reference exists (the build container); skipped on the GPU box, which has no /root/reference.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from flac_writer import encode_flac_quick

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE_TESTS = "/root/reference/tests/tests.py"

RUNNER = r"""
import sys, unittest, importlib.util
import pandas as pd
from voicemap_b200 import audio_io
sys.modules['soundfile'] = audio_io                       # tests.py: `import soundfile as sf`; sf.read(path)
if not hasattr(pd, 'value_counts'):
    pd.value_counts = lambda values: pd.Series(values).value_counts()
spec = importlib.util.spec_from_file_location('reference_tests', sys.argv[1])
module = importlib.util.module_from_spec(spec)
spec.loader.exec_module(module)
import voicemap.librispeech, voicemap.utils
assert voicemap.librispeech.LibriSpeechDataset.__module__ == 'voicemap_b200.librispeech'
suite = unittest.defaultTestLoader.loadTestsFromModule(module)
result = unittest.TextTestRunner(verbosity=2).run(suite)
print('RAN', result.testsRun, 'FAILED', len(result.failures) + len(result.errors))
sys.exit(0 if result.wasSuccessful() and result.testsRun == 3 else 1)
"""


def _write_corpus(root):
    rng = np.random.default_rng(0)
    lines = ["; miniature LibriSpeech for the reference's tests", ";ID  |SEX| SUBSET           |MINUTES| NAME"]
    speakers = [84] + list(range(200, 220))
    for s, speaker in enumerate(speakers):
        chapter = 121123 if speaker == 84 else 1000 + s
        folder = os.path.join(root, "data", "LibriSpeech", "dev-clean", str(speaker), str(chapter))
        os.makedirs(folder)
        lines.append("{:<5}| {} | dev-clean        | 8.{:02d} | Reader {}".format(speaker, "FM"[s % 2], s, s))
        for u in range(6):
            samples = 48000 + 500 * (u + 1) + 7 * s
            if speaker == 84 and u == 0:      # the file tests.py:74 opens: noise without DC offset, like speech (the
                # reference scales by the RMS of the un-centred batch, so its RMS assertion presumes zero-mean audio)
                pcm = np.round(2500 * rng.standard_normal(samples)).astype(np.int64)
                pcm -= int(np.round(pcm.mean()))
                data = encode_flac_quick(pcm)
            else:                             # content is irrelevant to the batcher tests: constant-coded, tiny
                data = encode_flac_quick(None, constant=(37 * (s + 1), samples))
            with open(os.path.join(folder, "{}-{}-{:04d}.flac".format(speaker, chapter, u)), "wb") as handle:
                handle.write(data)
    with open(os.path.join(root, "data", "LibriSpeech", "SPEAKERS.TXT"), "w") as handle:
        handle.write("\n".join(lines) + "\n")


@pytest.mark.skipif(not os.path.exists(REFERENCE_TESTS), reason="the reference tree is only present in the build container")
def test_reference_test_file_passes_against_the_drop_in(tmp_path):
    _write_corpus(str(tmp_path))
    env = dict(os.environ, VOICEMAP_PATH=str(tmp_path), PYTHONPATH=ROOT)
    run = subprocess.run([sys.executable, "-W", "ignore", "-c", RUNNER, REFERENCE_TESTS], cwd=ROOT, env=env,
                         capture_output=True, text=True, timeout=600)
    report = run.stdout + run.stderr
    assert run.returncode == 0, report
    assert "RAN 3 FAILED 0" in report, report
    for name in ("test_verification_batch", "test_n_shot_task", "test_whitening_no_batch"):
        assert name in report and "{} ".format(name) in report, report
    assert os.path.exists(os.path.join(str(tmp_path), "data", "dev-clean.index.csv"))   # indexed from FLAC headers


@pytest.mark.skipif(not os.path.exists(REFERENCE_TESTS), reason="the reference tree is only present in the build container")
def test_items_equal_the_reference_batchers_items(tmp_path):
    """voicemap/librispeech.py itself, imported unchanged, against ours on the same FLAC tree: identical index tables,
    identical clips and labels for every item in every (stochastic, pad, label) mode, and identical consumption of the
    numpy random stream -- including through our fragment-only, threaded batch path; and, seed by seed, the very same
    alike / differing pairs and k-way n-shot tasks (our array-based draws make the calls on numpy's global RandomState
    that pandas' DataFrame.sample makes for the reference)."""
    _write_corpus(str(tmp_path))
    script = os.path.join(ROOT, "tests", "golden", "run_reference_batcher.py")
    run = subprocess.run([sys.executable, "-W", "ignore", script, str(tmp_path)], cwd=str(tmp_path),
                         env=dict(os.environ, VOICEMAP_PATH=str(tmp_path), PYTHONPATH=ROOT),
                         capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, run.stdout + run.stderr
    verdict = json.loads(run.stdout.strip().splitlines()[-1])
    assert verdict["mismatches"] == [] and verdict["items"] == 5 * 126 and verdict["draws"] == 40, verdict
