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


def _write_irregular_corpus(root):
    """30 speakers with 1 - 7 utterances of 1 - 6 s each: files shorter than the fragment, speakers with a single
    usable file, speakers with none."""
    rng = np.random.default_rng(11)
    lines = ["; irregular miniature LibriSpeech", ";ID  |SEX| SUBSET           |MINUTES| NAME"]
    for s in range(30):
        speaker, chapter = 700 + s, 40 + s
        folder = os.path.join(root, "data", "LibriSpeech", "dev-clean", str(speaker), str(chapter))
        os.makedirs(folder)
        lines.append("{:<5}| {} | dev-clean        | 9.{:02d} | Reader {}".format(speaker, "MF"[s % 2], s, s))
        for u in range(int(rng.integers(1, 8))):
            samples = int(rng.integers(16000, 96000))
            with open(os.path.join(folder, "{}-{}-{:04d}.flac".format(speaker, chapter, u)), "wb") as handle:
                handle.write(encode_flac_quick(None, constant=(13 * (s + 1), samples)))
    with open(os.path.join(root, "data", "LibriSpeech", "SPEAKERS.TXT"), "w") as handle:
        handle.write("\n".join(lines) + "\n")


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
    assert verdict["reduced"] == 20 and verdict.get("skipped", 0) == 0 and verdict.get("both_failed", 0) == 0, verdict


@pytest.mark.skipif(not os.path.exists(REFERENCE_TESTS), reason="the reference tree is only present in the build container")
def test_items_and_draws_equal_the_reference_on_an_irregular_corpus(tmp_path):
    """Same side-by-side run on a ragged tree (files shorter than the fragment, speakers with one usable file or none):
    items in all modes, pair draws and n-shot tasks agree -- including the draws neither batcher can make."""
    _write_irregular_corpus(str(tmp_path))
    script = os.path.join(ROOT, "tests", "golden", "run_reference_batcher.py")
    run = subprocess.run([sys.executable, "-W", "ignore", script, str(tmp_path)], cwd=str(tmp_path),
                         env=dict(os.environ, VOICEMAP_PATH=str(tmp_path), PYTHONPATH=ROOT),
                         capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, run.stdout + run.stderr
    verdict = json.loads(run.stdout.strip().splitlines()[-1])
    assert verdict["mismatches"] == [] and verdict["items"] > 300 and verdict["draws"] == 40, verdict
    assert verdict["reduced"] == 20 and verdict.get("both_failed", 0) > 0, verdict


SCRIPT_RUNNER = r"""
import json, re, sys
import numpy as np
import voicemap_b200.models as M

record = {}


def fit_generator(self, generator=None, steps_per_epoch=None, epochs=1, validation_data=None, validation_steps=None,
                  workers=1, use_multiprocessing=False, callbacks=None, **extra):
    # training itself needs the device; everything the script does before it, and what it hands over, is checked here
    (left, right), labels = next(generator)
    (vleft, vright), vlabels = next(validation_data)
    record["calls"] = record.get("calls", 0) + 1
    record.setdefault("filters", []).append(self.layers[2].filters)
    record.update(model=type(self).__name__, loss=self.loss, optimizer=type(self.optimizer).__name__,
                  clipnorm=self.optimizer.clipnorm, metrics=self.metrics, steps_per_epoch=steps_per_epoch, epochs=epochs,
                  validation_steps=validation_steps, workers=workers, use_multiprocessing=use_multiprocessing,
                  callbacks=[type(c).__name__ for c in callbacks], extra=sorted(extra),
                  batch=[list(left.shape), list(right.shape), list(labels.shape)], dtype=str(left.dtype),
                  labels=labels[:, 0].tolist(), validation_batch=list(vleft.shape),
                  rms=float(np.sqrt(np.mean(np.square(left)))), encoder=type(self.layers[2]).__name__,
                  input_shape=list(self.input_shape), params=self.count_params())
    return []


M._ModelBase.fit_generator = fit_generator
path = sys.argv[1]
source = re.sub(r"^(\s*)print (.+)$", r"\1print(\2)", open(path).read(), flags=re.M)   # the script's one py2 statement
exec(compile(source, path, "exec"), {"__name__": "__main__", "__file__": path})
print("RECORD " + json.dumps(record))
"""


def _tiny_subset(root, subset, speakers, lines, base=40000, files=3):
    for s in speakers:
        folder = os.path.join(root, "data", "LibriSpeech", subset, str(s), "1")
        os.makedirs(folder)
        lines.append("{:<5}| {} | {:<16} | 30.00 | Reader {}".format(s, "FM"[s % 2], subset, s))
        for u in range(files):
            samples = base + 4000 * (u % 3) + s % 97                  # 2.5 - 3.0 s by default: padded to 3 s by pad=True
            with open(os.path.join(folder, "{}-1-{:04d}.flac".format(s, u)), "wb") as handle:
                handle.write(encode_flac_quick(None, constant=(11 * (s % 50 + 1), samples)))


@pytest.fixture(scope="module")
def three_subset_corpus(tmp_path_factory):
    """train-clean-100 / train-clean-360 (2 x 420 speakers, wide_vs_tall.py samples up to 800 of them) and dev-clean
    (40 speakers, files longer than 3 s because some scripts build the validation set without padding); every file
    is a genuine, constant-coded FLAC stream of ~170 bytes."""
    root = str(tmp_path_factory.mktemp("librispeech"))
    lines = ["; miniature corpus", ";ID  |SEX| SUBSET           |MINUTES| NAME"]
    _tiny_subset(root, "train-clean-100", range(1000, 1420), lines)
    _tiny_subset(root, "train-clean-360", range(3000, 3420), lines)
    # 5 files per dev speaker: a 32-pair differing draw may cover 32 of the 40 speakers and must still find 32 files of
    # the other 8 (with 3 files it failed once in a few runs: "Cannot take a larger sample than population")
    _tiny_subset(root, "dev-clean", range(500, 540), lines, base=49000, files=5)
    with open(os.path.join(root, "data", "LibriSpeech", "SPEAKERS.TXT"), "w") as handle:
        handle.write("\n".join(lines) + "\n")
    return root


def _run_reference_script(name, root):
    env = dict(os.environ, VOICEMAP_PATH=root, PYTHONPATH=os.pathsep.join([ROOT, os.path.join(ROOT, "compat")]))
    run = subprocess.run([sys.executable, "-W", "ignore", "-c", SCRIPT_RUNNER,
                          "/root/reference/experiments/{}.py".format(name)],
                         cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert run.returncode == 0, run.stdout[-3000:] + run.stderr[-3000:]
    line = [text for text in run.stdout.splitlines() if text.startswith("RECORD ")][-1]
    return json.loads(line[len("RECORD "):]), run.stdout


needs_reference = pytest.mark.skipif(not os.path.exists(REFERENCE_TESTS),
                                     reason="the reference tree is only present in the build container")


@needs_reference
def test_reference_train_siamese_script_runs_unchanged_up_to_training(three_subset_corpus):
    """experiments/train_siamese.py of the reference, executed from where it lies with its import lines untouched
    (`keras.*` served by compat/keras, `voicemap.*` and `config` by this repo; its Python-2 `print` statements are
    rewritten in memory): datasets over a miniature three-subset LibriSpeech tree, the batch generators, the two
    builders, compile, summary, plot_model and the four callbacks all go through, and `fit_generator` -- intercepted,
    because training needs the device -- receives what the reference passes to Keras."""
    record, stdout = _run_reference_script("train_siamese", three_subset_corpus)
    assert record["model"] == "SiameseModel" and record["encoder"] == "EncoderModel" and record["calls"] == 1
    assert record["loss"] == "binary_crossentropy" and record["optimizer"] == "Adam" and record["clipnorm"] == 1.0
    assert record["metrics"] == ["accuracy"] and record["input_shape"] == [12000, 1]
    assert record["params"] == 1023808 + 2 + 2560                       # encoder + Dense(1) + BN statistics
    assert (record["steps_per_epoch"], record["epochs"], record["validation_steps"]) == (500, 50, 100)
    assert record["use_multiprocessing"] is True and record["workers"] >= 1 and record["extra"] == []
    assert record["callbacks"] == ["NShotEvaluationCallback", "CSVLogger", "ModelCheckpoint", "ReduceLROnPlateau"]
    assert record["batch"] == [[64, 12000, 1], [64, 12000, 1], [64, 1]] and record["dtype"] == "float64"
    assert record["labels"] == [0.0] * 32 + [1.0] * 32 and record["validation_batch"] == [64, 12000, 1]
    assert "Total params: 1,026,370" in stdout and "Trainable params: 1,023,810" in stdout   # print siamese.summary()


@needs_reference
@pytest.mark.parametrize("name,loss,calls", [("siamese_contrastive_loss", "contrastive_loss", 1),
                                             ("determine_variance", "binary_crossentropy", 10),
                                             ("n_seconds_accuracy", "binary_crossentropy", None),
                                             ("grid_search_siamese_network", "binary_crossentropy", None),
                                             ("wide_vs_tall", "binary_crossentropy", 10)])
def test_other_reference_experiments_reach_training(three_subset_corpus, name, loss, calls):
    """The reference's other training experiments, same arrangement: every one of them gets as far as each of its
    `fit_generator` calls with a well-formed first batch.  `wide_vs_tall.py` additionally replaces `dataset.df` by a
    reduced frame (fewer speakers) and must then be served batches from those speakers only."""
    record, _ = _run_reference_script(name, three_subset_corpus)
    assert record["model"] == "SiameseModel" and record["optimizer"] == "Adam"
    assert record["loss"] == (loss if loss != "contrastive_loss" else record["loss"])
    if loss == "contrastive_loss":
        assert "contrastive_loss" in str(record["loss"])
    if calls is not None:
        assert record["calls"] == calls
    if name == "grid_search_siamese_network":
        assert record["calls"] == 40 and sorted(set(record["filters"])) == [16, 32, 64, 128]
    if name == "n_seconds_accuracy":
        assert record["batch"][0][1] > 12000                         # the sweep ends on clips longer than 3 s
    assert record["calls"] >= 1 and record["batch"][2] == [record["batch"][0][0], 1]
    assert record["batch"][0] == record["batch"][1] and record["batch"][0][2] == 1
