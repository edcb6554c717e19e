"""Multi-process batch production behind fit_generator(workers=N, use_multiprocessing=True) -- voicemap_b200/prefetch.py.
CPU only: the producers never touch the device, and the training loop is exercised with a stub engine."""
import os
import time

import numpy as np
import pytest

from voicemap_b200 import prefetch


def _siamese_batches(n_batches=None, clips=6, samples=500, fail_at=None):
    """([x1, x2], y) batches with y derived from the inputs, so that a torn or recycled-too-early slot shows."""
    produced = 0
    while n_batches is None or produced < n_batches:
        if fail_at is not None and produced == fail_at:
            raise ValueError("corpus went away")
        x1 = np.random.standard_normal((clips, samples, 1))
        x2 = np.random.standard_normal((clips, samples, 1)).astype(np.float32)
        y = (x1.sum(axis=(1, 2)) + x2.sum(axis=(1, 2), dtype=np.float64))[:, None]
        produced += 1
        yield [x1, x2], y


def _check(batch):
    (x1, x2), y = batch
    assert isinstance(batch, tuple) and isinstance(batch[0], list)          # nesting survives
    assert x1.dtype == np.float64 and x2.dtype == np.float32 and x1.shape == (6, 500, 1)
    np.testing.assert_allclose(y[:, 0], x1.sum(axis=(1, 2)) + x2.sum(axis=(1, 2), dtype=np.float64), rtol=1e-12)
    return float(x1[0, 0, 0])


def test_flatten_layout_rebuild_roundtrip():
    batch = ([np.arange(6.0).reshape(2, 3), np.ones((4, 1), np.float32)], np.zeros((2, 1)), None, 3)
    arrays = []
    tree = prefetch._flatten(batch, arrays)
    entries, total = prefetch._layout(arrays)
    assert [e[0] for e in entries] == [0, 64, 128] and total == 192
    again = prefetch._rebuild(tree, arrays)
    assert isinstance(again, tuple) and isinstance(again[0], list) and again[2] is None and again[3] == 3
    assert again[0][0] is arrays[0]


def test_generator_producers_distinct_streams_and_slot_lifetime():
    np.random.seed(7)
    pre = prefetch.ProcessPrefetcher(_siamese_batches(), workers=3)
    try:
        seen = []
        previous = None
        for _ in range(25):
            batch = pre.next()
            if previous is not None:
                _check(previous)          # still intact: a batch stays valid through the next next() call
            seen.append(_check(batch))
            previous = None if len(seen) == 1 else batch      # (the very first batch is the parent's own, not a slot)
        assert len(set(seen)) == 25       # reseeded children: no two workers replay the same "random" batches
        assert len(pre.pool.blocks) == 8 and all(p.is_alive() for p in pre.pool.procs)   # 2 * workers + 2 slots
    finally:
        pre.close()
    assert not any(p.is_alive() for p in pre.pool.procs)
    for block in pre.pool.blocks:          # shared memory is unlinked
        assert not os.path.exists("/dev/shm/" + block.name.lstrip("/"))


def test_finite_generator_ends_and_errors_surface():
    pre = prefetch.ProcessPrefetcher(_siamese_batches(n_batches=5), workers=2)
    count = 0
    with pytest.raises(StopIteration):
        while True:
            _check(pre.next())
            count += 1
    pre.close()
    assert count == 1 + 2 * 4              # the parent's first batch, then each forked copy runs out on its own (as in Keras)

    pre = prefetch.ProcessPrefetcher(_siamese_batches(fail_at=3), workers=2)
    with pytest.raises(RuntimeError, match="corpus went away"):
        for _ in range(20):
            pre.next()
    assert pre.pool.closed


class _SlowSequence:
    """Deterministic batches whose cost varies, so that children finish out of order."""

    def __init__(self, length=23):
        self.length = length
        self.epoch = 0

    def __len__(self):
        return self.length

    def __getitem__(self, i):
        time.sleep(0.002 * ((i * 7) % 5))
        x = np.full((4, 100, 1), float(i + 1000 * self.epoch))
        return x, np.full((4, 1), float(i))

    def on_epoch_end(self):
        self.epoch += 1


def test_sequence_batches_come_back_in_order_and_follow_epoch_changes():
    seq = _SlowSequence()
    for epoch in range(2):
        indices = [s % len(seq) for s in range(30)]            # steps_per_epoch > len(sequence) wraps around
        pre = prefetch.SequencePrefetcher(seq, indices, workers=4, slots=5)
        got = []
        for _ in indices:
            x, y = pre.next()
            assert x.shape == (4, 100, 1) and float(x[0, 0, 0]) == y[0, 0] + 1000 * epoch
            got.append(int(y[0, 0]))
        with pytest.raises(StopIteration):
            pre.next()
        pre.close()
        assert got == indices
        seq.on_epoch_end()                                     # the next epoch's children are forked after this
    single = prefetch.SequencePrefetcher(seq, [3], workers=4)   # one step: nothing to fork
    assert single.pool is None and int(single.next()[1][0, 0]) == 3
    single.close()


def test_fit_generator_uses_the_producers(monkeypatch):
    """The training loop with workers=4, use_multiprocessing=True, on a stub engine: generator and Sequence inputs."""
    import torch
    from voicemap_b200 import training
    from voicemap_b200.keras_compat import Adam
    from voicemap_b200.models import build_siamese_net, get_baseline_convolutional_encoder

    steps = []

    class StubEngine:
        kind = "siamese"
        bwd_precision = 1

        def __init__(self, model, optimizer, loss):
            self.optimizer, self.loss = optimizer, loss

        def set_sync_bn(self, allreduce, world, peers=None):
            pass

        def set_gradient_buckets(self, enabled=True):
            pass

        def siamese_step(self, x1, x2, y, allreduce=None, world=1):
            _check(([x1, x2], y))
            steps.append(float(y[0, 0]))
            return torch.tensor(float(np.mean(y))), torch.tensor(1.0)

        def sync_to_model(self):
            pass

    monkeypatch.setattr(training, "TrainEngine", StubEngine)
    monkeypatch.setattr(training, "_producer_processes", lambda workers, multi: 3 if multi and workers > 1 else 0)
    model = build_siamese_net(get_baseline_convolutional_encoder(16, 8), (500, 1))
    model.compile(loss="binary_crossentropy", optimizer=Adam(), metrics=["accuracy"])
    history = model.fit_generator(_siamese_batches(), steps_per_epoch=7, epochs=2, verbose=0, workers=4,
                                  use_multiprocessing=True)
    assert len(history) == 2 and len(steps) == 14 and len(set(steps)) == 14

    class Seq:
        def __len__(self):
            return 5

        def __getitem__(self, i):
            rng = np.random.default_rng(i)
            x1, x2 = rng.standard_normal((6, 500, 1)), rng.standard_normal((6, 500, 1)).astype(np.float32)
            return [x1, x2], (x1.sum(axis=(1, 2)) + x2.sum(axis=(1, 2), dtype=np.float64))[:, None]

        def on_epoch_end(self):
            pass

    # default arguments: the single background thread (generator) / in-line indexing (Sequence), same loop
    del steps[:]
    model.fit_generator(_siamese_batches(), steps_per_epoch=3, epochs=2, verbose=0)
    model.fit_generator(Seq(), epochs=1, verbose=0)
    assert len(steps) == 6 + 5

    # a failing step closes the producers on its way out
    def boom(self, x1, x2, y, allreduce=None, world=1):
        raise RuntimeError("device lost")
    monkeypatch.setattr(StubEngine, "siamese_step", boom)
    model._trainer = None
    import multiprocessing
    before = len(multiprocessing.active_children())
    with pytest.raises(RuntimeError, match="device lost"):
        model.fit_generator(_siamese_batches(), steps_per_epoch=3, epochs=1, verbose=0, workers=4, use_multiprocessing=True)
    time.sleep(0.2)
    assert len(multiprocessing.active_children()) <= before
    monkeypatch.undo()
    monkeypatch.setattr(training, "TrainEngine", StubEngine)
    monkeypatch.setattr(training, "_producer_processes", lambda workers, multi: 3 if multi and workers > 1 else 0)
    model._trainer = None

    del steps[:]
    model.fit_generator(Seq(), epochs=2, verbose=0, workers=4, use_multiprocessing=True)
    assert len(steps) == 10 and steps[:5] == steps[5:]          # index order, both epochs
    assert training._producer_processes.__name__ == "<lambda>"


def test_producer_count_policy():
    from voicemap_b200.training import _producer_processes
    assert _producer_processes(1, True) == 0 and _producer_processes(8, False) == 0 and _producer_processes(None, True) == 0
    cores = len(os.sched_getaffinity(0))
    if cores > 1:
        assert _producer_processes(64, True) == min(64, cores, 16)


def test_producers_run_the_real_batcher_on_flac_files(tmp_path):
    """Forked producers running the reference-style generator expression over LibriSpeechDataset: FLAC files on disk,
    the decoder library's own threads inside each child, preprocessing, shared-memory hand-over."""
    from flac_writer import encode_flac_quick
    from voicemap_b200 import utils
    from voicemap_b200.librispeech import LibriSpeechDataset
    root = tmp_path / "data" / "LibriSpeech"
    lines = ["; tiny"]
    rng = np.random.default_rng(0)
    for s in range(40):
        folder = root / "dev-clean" / str(100 + s) / "1"
        folder.mkdir(parents=True)
        lines.append("{} | {} | dev-clean | 9.0 | R{}".format(100 + s, "FM"[s % 2], s))
        for u in range(3):
            pcm = np.round(1000 * rng.standard_normal(17000 + 100 * u) + 40 * s).astype(np.int64)
            (folder / "{}-1-{:04d}.flac".format(100 + s, u)).write_bytes(encode_flac_quick(pcm))
    (root / "SPEAKERS.TXT").write_text("\n".join(lines) + "\n")
    dataset = LibriSpeechDataset("dev-clean", 1, data_path=str(tmp_path), cache=False)
    pre = utils.BatchPreProcessor("siamese", utils.preprocess_instances(4))
    batches = (pre(batch) for batch in dataset.yield_verification_batches(16))
    producers = prefetch.ProcessPrefetcher(batches, workers=2)
    try:
        firsts = []
        for _ in range(7):
            (left, right), labels = producers.next()
            assert left.shape == right.shape == (16, 4000, 1) and left.dtype == np.float64
            assert labels[:, 0].tolist() == [0.0] * 8 + [1.0] * 8
            # whitened: the gain comes from the un-centred batch (reference quirk), so DC offsets leave the RMS below 0.038
            assert 0.01 < float(np.sqrt(np.mean(np.square(left)))) <= 0.038021 + 1e-9
            assert abs(float(left.mean(axis=1).max())) < 1e-12
            firsts.append(float(left[0, 0, 0]))
        assert len(set(firsts)) == 7
    finally:
        producers.close()


def test_batch_kept_past_close_stays_mapped():
    """A batch handed out by a prefetcher is a view of a shared-memory slot; closing the pool (end of an epoch, an
    error, fit's end) must not unmap it under a caller who still holds it -- the block lingers and is released when the
    view is gone."""
    np.random.seed(11)
    pre = prefetch.ProcessPrefetcher(_siamese_batches(), workers=2)
    pre.next()                                   # the parent's own first batch
    kept = pre.next()                            # a slot view
    snapshot = [a.copy() for a in kept[0]] + [kept[1].copy()]
    blocks = list(pre.pool.blocks)
    pre.close()
    assert any(b in prefetch._LINGERING for b in blocks)            # deferred, not unmapped
    _check(kept)                                                    # would segfault on a dangling mapping
    np.testing.assert_array_equal(kept[0][0], snapshot[0])
    np.testing.assert_array_equal(kept[1], snapshot[2])
    del kept
    prefetch._release([])                                           # the next close retries the lingering blocks
    assert not any(b in prefetch._LINGERING for b in blocks)
