"""Host-side logic that needs no GPU: builder API surface, Keras-method subset, preprocessing, batcher sampling,
callbacks (mirrors the reference's tests/tests.py where it has tests: verification batches, n-shot tasks,
whitening)."""
import numpy as np
import pandas as pd
import pytest

from oracle import voicemap_oracle as O
from voicemap_b200 import utils
from voicemap_b200.keras_compat import (Adam, CSVLogger, Dense, ModelCheckpoint, ReduceLROnPlateau, clone_model,
                                        to_categorical)
from voicemap_b200.librispeech import LibriSpeechDataset
from voicemap_b200.models import (build_siamese_net, get_baseline_convolutional_encoder, load_model)


# ----------------------------------------------------------------------------------------------- builders
def test_encoder_builder_signature_and_layers():
    enc = get_baseline_convolutional_encoder(128, 64, (12000, 1))
    names = [l.name for l in enc.layers]
    assert names[:4] == ["conv1d_1", "batch_normalization_1", "spatial_dropout1d_1", "max_pooling1d_1"]
    assert names[-2:] == ["global_max_pooling1d_1", "dense_1"]
    assert enc.layers[3].config["pool_size"] == 4 and enc.layers[7].config["pool_size"] == 2
    assert enc.dropout == 0.05  # default of the reference signature
    w = enc.get_weights()
    assert [a.shape for a in w[:6]] == [(32, 1, 128), (128,), (128,), (128,), (128,), (128,)]
    assert w[-2].shape == (512, 64) and w[-1].shape == (64,)
    trainable = sum(a.size for i, a in enumerate(w) if i % 6 not in (4, 5) or i >= 24)
    assert trainable == 1_023_808


def test_classifier_add_pop_clone_roundtrip():
    clf = get_baseline_convolutional_encoder(16, 8, (4000, 1))
    clf.add(Dense(11, activation="softmax"))
    assert clf.layers[-1].name == "dense_2" and len(clf.get_weights()) == 28
    enc = clone_model(clf)
    enc.set_weights(clf.get_weights())
    enc.pop()
    assert len(enc.get_weights()) == 26 and enc.layers[-1].name == "dense_1"
    for a, b in zip(enc.get_weights(), clf.get_weights()[:26]):
        np.testing.assert_array_equal(a, b)
    with pytest.raises(ValueError):
        enc.set_weights(clf.get_weights())


def test_siamese_builder_contract():
    enc = get_baseline_convolutional_encoder(16, 8, dropout=0.0)
    sia = build_siamese_net(enc, (4000, 1), distance_metric="uniform_euclidean")
    assert sia.layers[2] is enc  # voicemap/utils.py:141
    assert [l.name for l in sia.layers if l is not enc] == ["input_1", "input_2", "subtract_embeddings",
                                                           "euclidean_distance", "dense_2"]
    assert sia.head_weights["head_kernel"].shape == (1, 1)
    l1 = build_siamese_net(get_baseline_convolutional_encoder(16, 8), (4000, 1), "weighted_l1")
    assert l1.head_weights["head_kernel"].shape == (8, 1)
    with pytest.raises(NotImplementedError):
        build_siamese_net(enc, (4000, 1), "cosine_distance")
    with pytest.raises(AssertionError):
        build_siamese_net(enc, (4000, 1), "manhattan")


def test_compile_accepts_reference_losses():
    sia = build_siamese_net(get_baseline_convolutional_encoder(16, 8), (4000, 1))
    sia.compile(loss="binary_crossentropy", optimizer=Adam(clipnorm=1.), metrics=["accuracy"])
    assert sia.loss == "binary_crossentropy" and sia.optimizer.clipnorm == 1.0 and sia.optimizer.epsilon == 1e-7
    sia.compile(loss=utils.contrastive_loss, optimizer=Adam(clipnorm=1., decay=2e-5))
    assert sia.loss == "contrastive_loss"
    with pytest.raises(NotImplementedError):
        sia.compile(loss="mse", optimizer=Adam())


def test_save_load_roundtrip(tmp_path):
    enc = get_baseline_convolutional_encoder(16, 8, dropout=0.0)
    sia = build_siamese_net(enc, (4000, 1))
    path = str(tmp_path / "m.npz")
    sia.save(path)
    back = load_model(path)
    for a, b in zip(sia.get_weights(), back.get_weights()):
        np.testing.assert_array_equal(a, b)
    assert back.layers[2].filters == 16


def test_predict_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    enc = get_baseline_convolutional_encoder(16, 8)
    with pytest.raises(Exception) as exc:
        enc.predict(np.zeros((1, 4000, 1), np.float32))
    assert "CUDA" in str(exc.value) or "cuda" in str(exc.value)


# ----------------------------------------------------------------------------------------------- preprocessing
def test_whiten_matches_oracle_and_reference_known_answer():
    rng = np.random.default_rng(0)
    clip = rng.normal(0, 0.3, 48000)
    batch = np.stack([clip, clip])[:, :, np.newaxis]  # reference tests/tests.py:76
    w = utils.whiten(batch)
    np.testing.assert_allclose(w, O.whiten_literal(batch), rtol=1e-12, atol=1e-15)
    assert abs(w[0].mean()) < 1e-9
    rms = np.sqrt(np.power(w[0, :, 0], 2).mean())
    assert abs(rms - 0.038021) < 0.038021 * 0.02  # the reference test allows places=5 on a zero-mean clip
    with pytest.raises(ValueError):
        utils.whiten(np.zeros((4, 10)))


def test_batch_preprocessor_modes():
    pre = utils.BatchPreProcessor("siamese", utils.preprocess_instances(4))
    a, b = np.random.rand(2, 64, 1), np.random.rand(2, 64, 1)
    [o1, o2], y = pre(([a, b], np.zeros((2, 1))))
    assert o1.shape == (2, 16, 1) and o2.shape == (2, 16, 1)
    clf = utils.BatchPreProcessor("classifier", utils.preprocess_instances(4), lambda y: y + 1)
    x, y = clf((a, np.zeros((2, 1))))
    assert x.shape == (2, 16, 1) and (y == 1).all()
    with pytest.raises(AssertionError):
        utils.BatchPreProcessor("other", None)


def test_contrastive_loss_numpy_matches_oracle():
    y = np.array([[0.], [1.], [1.]])
    p = np.array([[0.3], [0.2], [1.4]])
    assert np.isclose(utils.contrastive_loss(y, p), O.contrastive_loss(y, p))


def test_to_categorical():
    out = to_categorical(np.array([[0], [2], [1]]), 4)
    assert out.shape == (3, 4) and out[1, 2] == 1 and out.sum() == 3


# ----------------------------------------------------------------------------------------------- batcher
def _fake_dataset(seconds=1, n_speakers=8, files_per_speaker=5, pad=False, stochastic=True):
    rng = np.random.default_rng(0)
    rows, audio = [], {}
    for spk in range(n_speakers):
        for j in range(files_per_speaker):
            path = f"/fake/{spk}/{j}.flac"
            length = int(16000 * (1.2 + rng.random()))
            audio[path] = rng.normal(0, 0.1, length) + spk  # speaker id recoverable from the mean
            rows.append(dict(id=100 + spk, sex="M" if spk % 2 else "F", subset="dev-clean", minutes=10.0,
                             name=f"s{spk}", filepath=path, length=length, seconds=length / 16000.0))
    reader = lambda p: (audio[p], 16000)  # noqa: E731
    return LibriSpeechDataset("dev-clean", seconds, stochastic=stochastic, pad=pad, index=pd.DataFrame(rows),
                              reader=reader)


def test_verification_batch_like_reference_test():
    # reference tests/tests.py:16-29: alike pairs share a speaker, differing pairs do not; labels 0 then 1
    ds = _fake_dataset()
    [i1, i2], y = ds.build_verification_batch(8)
    assert i1.shape == (8, 16000, 1) and i2.shape == (8, 16000, 1) and y.shape == (8, 1)
    np.testing.assert_array_equal(y[:, 0], [0, 0, 0, 0, 1, 1, 1, 1])
    spk1, spk2 = np.round(i1.mean(axis=(1, 2))), np.round(i2.mean(axis=(1, 2)))
    assert (spk1[:4] == spk2[:4]).all() and (spk1[4:] != spk2[4:]).all()


def test_n_shot_task_like_reference_test():
    # reference tests/tests.py:31-68
    ds = _fake_dataset()
    for k, n in ((5, 1), (3, 2)):
        (q, q_label), (support, labels) = ds.build_n_shot_task(k, n)
        assert support.shape == (k * n, 16000) and len(labels) == k * n
        assert (labels[:n] == q_label).all()              # the first n support samples are the query speaker
        assert len(np.unique(labels)) == k                 # k unique speakers
        for c in range(k):
            assert len(np.unique(labels[c * n:(c + 1) * n])) == 1   # ordered [c1]*n + [c2]*n + ...
    with pytest.raises(ValueError):
        ds.build_n_shot_task(8)
    with pytest.raises(ValueError):
        ds.build_n_shot_task(1)


def test_getitem_padding_and_labels():
    ds = _fake_dataset(seconds=3, pad=True, stochastic=False)
    x, label = ds[0]
    assert x.shape == (48000,) and label == 100
    assert (x[int(16000 * 2.3):] == 0).all()  # deterministic padding appends zeros
    sex = LibriSpeechDataset("dev-clean", 1, label="sex", index=_fake_dataset().df.rename(
        columns={"speaker_id": "id", "speaker_minutes": "minutes"}).drop(columns=["id"], errors="ignore").assign(
        id=lambda d: d.index), reader=lambda p: (np.zeros(20000), 16000))
    assert sex[0][1] in (True, False)
    assert ds.num_classes() == 8 and len(ds) == 40


# ----------------------------------------------------------------------------------------------- n-shot eval + callbacks
class _StubEncoder:
    """Embeds a clip as (mean, 0): speakers are separable, so every task must be solved."""
    def predict(self, x):
        m = x.mean(axis=(1, 2))
        return np.stack([m, np.zeros_like(m)], axis=1) * 1000.0


class _StubSiamese:
    layers = [None, None, _StubEncoder()]

    def predict(self, x):
        a, b = _StubEncoder().predict(x[0]), _StubEncoder().predict(x[1])
        return np.linalg.norm(a - b, axis=1, keepdims=True)


def test_n_shot_task_evaluation_paths():
    ds = _fake_dataset()
    pre = utils.BatchPreProcessor("siamese", utils.preprocess_instances(4, whitening=False))
    assert utils.n_shot_task_evaluation(_StubSiamese(), ds, pre, num_tasks=6, n=1, k=4) == 6
    for dist in ("euclidean", "cosine", "dot_product"):
        got = utils.n_shot_task_evaluation(_StubSiamese(), ds, pre, num_tasks=4, n=2, k=3, distance=dist)
        assert 0 <= got <= 4
    assert utils.n_shot_task_evaluation(_StubSiamese(), ds, pre, 4, n=2, k=3, distance="euclidean") == 4


def test_nshot_callback_writes_logs_and_checkpoint_uses_them(tmp_path):
    ds = _fake_dataset()
    pre = utils.BatchPreProcessor("siamese", utils.preprocess_instances(4, whitening=False))
    cb = utils.NShotEvaluationCallback(4, 1, 3, ds, preprocessor=pre)
    cb.set_model(_StubSiamese())
    logs = {"loss": 1.0}
    cb.on_epoch_end(0, logs)
    assert logs["val_1-shot_acc"] == 1.0

    saved = []

    class M:
        optimizer = Adam(lr=0.5)
        def save(self, path):
            saved.append(path)
    ck = ModelCheckpoint(str(tmp_path / "m.npz"), monitor="val_1-shot_acc", mode="max", save_best_only=True)
    ck.set_model(M())
    ck.on_epoch_end(0, logs)
    ck.on_epoch_end(1, {"val_1-shot_acc": 0.5})
    assert len(saved) == 1
    rl = ReduceLROnPlateau(monitor="val_1-shot_acc", mode="max", patience=2)
    rl.set_model(M())
    for e in range(4):
        rl.on_epoch_end(e, {"val_1-shot_acc": 0.1})
    assert np.isclose(M.optimizer.lr, 0.05)
    log = CSVLogger(str(tmp_path / "log.csv"))
    log.on_train_begin()
    log.on_epoch_end(0, {"loss": 1.0, "val_1-shot_acc": 0.5})
    log.on_train_end()
    assert "val_1-shot_acc" in open(tmp_path / "log.csv").read()


def test_keras_hdf5_reader_on_synthetic_file(tmp_path):
    """The pure-Python HDF5 reader is exercised on the reference's real checkpoint when building the golden fixture
    (tests/golden/make_checkpoint_fixture.py); here: the committed fixture carries the survey's known answers."""
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "checkpoint_f32.npz"))
    k = z["w_conv1_kernel"]
    assert k.shape == (32, 1, 32)
    assert abs(float(k.mean()) - 0.000365) < 1e-6 and abs(float(k.std()) - 0.115049) < 1e-6
    assert abs(float(k.min()) + 0.36339) < 1e-5 and abs(float(k.max()) - 0.40413) < 1e-5
    assert abs(float(z["head_bias"][0]) + 2.692142) < 1e-6
    assert z["w_dense_kernel"].shape == (128, 128) and z["head_kernel"].shape == (128, 1)
    assert z["w_bn1_var"].min() < 1e-6 and z["w_bn4_var"].max() > 29


def test_load_model_selects_the_checkpoints_own_architecture():
    """The checkpoint the reference ships (models/n_seconds/...) was trained with four MaxPooling1D(2)
    and a weighted_l1 head; load_model must rebuild exactly that (first_pool = 2), not force the weights into the
    current architecture.  The file exists only where the reference is mounted."""
    import os
    import pytest
    path = "/root/reference/models/n_seconds/siamese__nseconds_3.0__filters_32__embed_64__drop_0.05__r_0.hdf5"
    if not os.path.exists(path):
        pytest.skip("reference checkpoint not available on this box")
    from voicemap_b200.models import load_model
    m = load_model(path)
    assert m.distance_metric == "weighted_l1"
    enc = m.layers[2]
    assert enc.first_pool == 2 and enc.filters == 32 and enc.embedding_dimension == 128
    assert [l.config["pool_size"] for l in enc.layers if l.class_name == "MaxPooling1D"] == [2, 2, 2, 2]
    clone = enc._clone()
    assert clone.first_pool == 2 and clone.get_config()["first_pool"] == 2


def test_first_pool_argument_is_validated():
    import pytest
    from voicemap_b200.models import get_baseline_convolutional_encoder
    assert get_baseline_convolutional_encoder(16, 8).first_pool == 4          # the reference architecture
    assert get_baseline_convolutional_encoder(16, 8, first_pool=2).first_pool == 2
    with pytest.raises(ValueError):
        get_baseline_convolutional_encoder(16, 8, first_pool=3)


def _randomize(model, seed):
    rng = np.random.default_rng(seed)
    ws = [rng.normal(size=w.shape).astype(np.float32) for w in model.get_weights()]
    model.set_weights(ws)
    return ws


def test_hdf5_checkpoints_round_trip_all_model_kinds(tmp_path):
    """model.save('*.hdf5') writes the Keras 2.2.x layout (root attrs, /model_weights/<layer>/<weight>); load_model
    restores architecture (first pool, dropout, head kind, input shape) and every weight bit for bit."""
    from voicemap_b200.keras_compat import Adam, Dense
    from voicemap_b200.keras_hdf5 import KerasH5, load_keras_weights
    from voicemap_b200.models import build_siamese_net, get_baseline_convolutional_encoder, load_model
    sia = build_siamese_net(get_baseline_convolutional_encoder(16, 8, dropout=0.1, first_pool=2), (4000, 1), "weighted_l1")
    sia.compile(loss="binary_crossentropy", optimizer=Adam(clipnorm=1.), metrics=["accuracy"])
    ws = _randomize(sia, 1)
    path = str(tmp_path / "siamese.hdf5")
    sia.save(path)
    m = load_model(path)
    assert m.distance_metric == "weighted_l1" and m.encoder.first_pool == 2 and m.encoder.dropout == 0.1
    assert tuple(m.input_shape) == (4000, 1)
    assert all(np.array_equal(a, b) for a, b in zip(ws, m.get_weights()))
    # same group / attribute structure as the reference's own checkpoint
    cfg, layers = load_keras_weights(path)
    assert [n for n, _ in layers] == ["input_1", "input_2", "sequential_1", "subtract_1", "lambda_1", "dense_2"]
    names = [w for w, _ in layers[2][1]]
    assert names[0] == "sequential_1/conv1d_1/kernel:0" and names[-1] == "sequential_1/batch_normalization_4/moving_variance:0"
    assert names.index("sequential_1/dense_1/bias:0") < names.index("sequential_1/batch_normalization_1/moving_mean:0")
    f = KerasH5(path)
    assert f.attr("/", "keras_version") == "2.2.2" and f.attr("/", "backend") == "tensorflow"
    assert '"clipnorm": 1.0' in f.attr("/", "training_config")
    assert cfg["config"]["layers"][2]["config"][3]["config"]["pool_size"] == [2]

    clf = get_baseline_convolutional_encoder(16, 8, (12000, 1))
    clf.add(Dense(10, activation="softmax"))
    ws = _randomize(clf, 2)
    path = str(tmp_path / "classifier.h5")
    clf.save(path)
    m = load_model(path)
    assert m._head == {"units": 10, "activation": "softmax"} and m.input_shape == (12000, 1) and m.first_pool == 4
    assert all(np.array_equal(a, b) for a, b in zip(ws, m.get_weights()))

    sia2 = build_siamese_net(get_baseline_convolutional_encoder(16, 8), (12000, 1))
    ws = _randomize(sia2, 3)
    path = str(tmp_path / "euclid.hdf5")
    sia2.save(path)
    m = load_model(path)
    assert m.distance_metric == "uniform_euclidean"
    assert all(np.array_equal(a, b) for a, b in zip(ws, m.get_weights()))
    # other suffixes keep the npz container
    sia2.save(str(tmp_path / "model.npz"))
    assert not open(tmp_path / "model.npz", "rb").read(8).startswith(b"\x89HDF")


def test_hdf5_writer_structure(tmp_path):
    """Byte-level invariants libhdf5 relies on: superblock fields, end-of-file address, 8-byte alignment of every
    object, symbol-table entries sorted by name, and the float32 datatype message identical to the one in the
    reference's checkpoint."""
    import struct
    from voicemap_b200.keras_hdf5 import UNDEF, KerasH5, save_keras_weights
    layers = [("zeta", [("zeta/kernel:0", np.arange(6, dtype=np.float32).reshape(2, 3)), ("zeta/bias:0", np.zeros(3, np.float32))]),
              ("alpha", []), ("mid", [("mid/w:0", np.ones((1,), np.float32))])]
    path = str(tmp_path / "w.h5")
    save_keras_weights(path, {"class_name": "Sequential", "config": []}, layers)
    b = open(path, "rb").read()
    assert b[:8] == b"\x89HDF\r\n\x1a\n" and b[8:16] == bytes([0, 0, 0, 0, 0, 8, 8, 0])
    base, free, eof, driver = struct.unpack_from("<QQQQ", b, 24)
    assert (base, free, eof, driver) == (0, UNDEF, len(b), UNDEF) and len(b) % 8 == 0
    f = KerasH5(path)
    seen = []

    def walk(addr):
        assert addr % 8 == 0 and b[addr] == 1          # version-1 object header, aligned
        kids = f.children(addr)
        seen.append(list(kids))
        for child in kids.values():
            walk(child)
    walk(f.root)
    assert seen[0] == ["model_weights"] and seen[1] == ["alpha", "mid", "zeta"]      # strcmp order inside the node
    assert f.attr("/model_weights", "layer_names") == ["zeta", "alpha", "mid"]       # attribute keeps model order
    assert f.attr("/model_weights/alpha", "weight_names") == []
    assert np.array_equal(f.dataset("/model_weights/zeta/zeta/kernel:0"), np.arange(6, dtype=np.float32).reshape(2, 3))
    ref_dtype = bytes.fromhex("11201f000400000000002000170800177f000000")   # float32 message of the reference's file
    for mtype, data in f._messages(f.resolve("/model_weights/mid/mid/w:0")):
        if mtype == 0x03:
            assert data[:20] == ref_dtype


def test_summary_of_both_model_kinds_counts_parameters():
    """model.summary() (experiments/train_siamese.py:58, train_classifier.py:116): one line per layer, the siamese net
    lists its shared encoder as a nested Sequential; totals as Keras reports them (1 023 808
    trainable encoder parameters + 2 560 moving statistics at filters=128, emb=64; +2 for the siamese head)."""
    from voicemap_b200.models import build_siamese_net, get_baseline_convolutional_encoder
    enc = get_baseline_convolutional_encoder(128, 64)
    lines = []
    enc.summary(print_fn=lines.append)
    text = "\n".join(lines)
    assert "conv1d_1 (Conv1D)" in text and "dense_1 (Dense)" in text
    assert f"Total params: {1023808 + 2560:,}" in text
    sia = build_siamese_net(enc, (12000, 1))
    lines = []
    sia.summary(print_fn=lines.append)
    text = "\n".join(lines)
    assert "sequential_1 (Sequential)" in text and "dense_2 (Dense)" in text
    assert f"Total params: {1023808 + 2560 + 2:,}" in text


def test_verification_batches_on_a_small_corpus():
    """The reference draws pairs with ``df.sample(n, weights='length')`` (voicemap/librispeech.py:143-167); current pandas
    rejects that draw on small corpora (n * max weight > total weight).  The batcher keeps the reference's semantics
    (distinct rows, probability proportional to length) through numpy.  (The reference's differing-pair draw needs
    more speakers than pairs: it excludes every speaker of the first draw.)"""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples"))
    from synthetic_speakers import SyntheticCorpus
    from voicemap_b200.librispeech import LibriSpeechDataset
    va = SyntheticCorpus(40, 3, subset="synthetic-dev", seed=1)
    valid = LibriSpeechDataset("synthetic-dev", 3, stochastic=False, pad=True, index=va.index, reader=va.reader)
    np.random.seed(0)
    for _ in range(5):
        (a, b), y = valid.build_verification_batch(32)
        assert a.shape == b.shape == (32, 48000, 1) and y.shape == (32, 1)
        assert y[:16].sum() == 0 and y[16:].sum() == 16
    pairs = valid.get_differing_pairs(16)
    spk = valid.df.set_index("id")["speaker_id"]
    assert all(spk[i] != spk[j] for i, j in pairs)
    assert all(spk[i] == spk[j] for i, j in valid.get_alike_pairs(16))


def test_empty_batches_give_empty_results_without_a_device():
    from voicemap_b200.models import build_siamese_net, get_baseline_convolutional_encoder
    enc = get_baseline_convolutional_encoder(16, 8)
    assert enc.predict(np.zeros((0, 4000, 1))).shape == (0, 8)
    sia = build_siamese_net(enc, (4000, 1))
    assert sia.predict([np.zeros((0, 4000, 1)), np.zeros((0, 4000, 1))]).shape == (0, 1)
    with pytest.raises(ValueError):
        sia.predict([np.zeros((0, 4000, 1)), np.zeros((0, 3000, 1))])


def test_corpus_indexing_from_disk_and_cache(tmp_path):
    """SPEAKERS.TXT parsing, the <subset>/<speaker>/<chapter>/*.flac walk, the per-subset CSV cache and the length
    filter (voicemap/librispeech.py:43-101, 243-281), on a miniature corpus written here."""
    from voicemap_b200.librispeech import LibriSpeechDataset, read_speaker_table
    root = tmp_path / "data" / "LibriSpeech"
    lengths = {}
    for spk, chapters in ((14, (100, 101)), (16, (200,)), (17, (300,))):
        for ch in chapters:
            d = root / "dev-clean" / str(spk) / str(ch)
            d.mkdir(parents=True)
            for u in range(3):
                p = d / f"{spk}-{ch}-{u:04d}.flac"
                p.write_bytes(b"")
                lengths[str(p)] = 16000 * (1 + u)        # 1 s, 2 s, 3 s
            (d / f"{spk}-{ch}.trans.txt").write_text("not audio")
    (root / "SPEAKERS.TXT").write_text(
        "; comment\n;ID  |SEX| SUBSET           |MINUTES| NAME\n"
        "14   | F | dev-clean  | 25.03 | Reader One\n"
        "16   | M | dev-clean  | 25.11 | Reader Two\n"
        "17   | M | dev-clean  | 25.04 | Reader | With | Pipes\n"      # malformed: dropped, like read_csv does
        "19   | F | train-clean-100  | 25.19 | Elsewhere\n")
    table = read_speaker_table(str(root / "SPEAKERS.TXT"))
    assert table["id"].tolist() == [14, 16, 19] and table["sex"].tolist() == ["F", "M", "F"]
    assert table["subset"].tolist() == ["dev-clean", "dev-clean", "train-clean-100"]

    calls = []

    def reader(path):
        calls.append(path)
        return np.zeros(lengths[path]), 16000

    ds = LibriSpeechDataset("dev-clean", 1.5, reader=reader, data_path=str(tmp_path))
    # speakers 14 (2 chapters) and 16 (1 chapter) survive the merge; of 9 files the 1 s ones are too short
    assert len(ds) == 6 and ds.num_classes() == 2 and ds.unique_speakers == 2
    assert set(ds.df.columns) >= {"speaker_id", "speaker_minutes", "sex", "subset", "name", "filepath", "length",
                                  "seconds", "id"}
    assert ds.df["id"].tolist() == list(range(6)) and set(ds.df["speaker_id"]) == {14, 16}
    assert (tmp_path / "data" / "dev-clean.index.csv").exists()
    decoded = len(calls)
    assert decoded == 12                                   # every .flac once (speaker 17's too), no .txt
    again = LibriSpeechDataset(["dev-clean"], 0.5, reader=reader, data_path=str(tmp_path))
    assert len(calls) == decoded and len(again) == 9       # served by the cache; shorter minimum keeps all files
    x, label = again[0]
    assert x.shape == (8000,) and label in (14, 16)


def test_pair_draws_follow_the_reference_distribution():
    """Alike pairs: uniform over (anchor, same-speaker file) combinations, the anchor itself included; differing pairs
    never share a speaker; draws prefer long files (voicemap/librispeech.py:143-167)."""
    ds = _fake_dataset(n_speakers=6, files_per_speaker=4)
    np.random.seed(3)
    spk = ds.df["speaker_id"].to_numpy()
    same_file = 0
    for _ in range(200):
        pairs = ds.get_alike_pairs(3)
        assert len(pairs) == 3 and all(spk[a] == spk[b] for a, b in pairs)
        same_file += sum(a == b for a, b in pairs)
        assert all(spk[a] != spk[b] for a, b in ds.get_differing_pairs(3))
    assert 0.15 < same_file / 600.0 < 0.35                 # 1 in 4 partners is the anchor itself
    ds._weight[:] = 1.0
    ds._weight[5] = 1e6
    assert all(5 in ds._draw(2) for _ in range(20))


def test_n_shot_evaluation_batched_equals_per_task_calls():
    ds = _fake_dataset()
    pre = utils.BatchPreProcessor("siamese", utils.preprocess_instances(4))
    for n, k, dist in ((2, 3, "euclidean"), (3, 2, "cosine"), (2, 4, "dot_product")):
        np.random.seed(11)
        one = utils.n_shot_task_evaluation(_StubSiamese(), ds, pre, 7, n, k, distance=dist)
        np.random.seed(11)
        many = utils.n_shot_task_evaluation(_StubSiamese(), ds, pre, 7, n, k, distance=dist, tasks_per_launch=3)
        assert one == many
    with pytest.raises(ValueError):
        utils.n_shot_task_evaluation(_StubSiamese(), ds, pre, 2, 2, 3, distance="manhattan")
    with pytest.raises(ValueError):
        utils.n_shot_task_evaluation(_StubSiamese(), ds, pre, 2, 2, 3, network_type="other")


def test_precision_is_validated_and_survives_clone():
    from voicemap_b200.keras_compat import clone_model
    from voicemap_b200.models import get_baseline_convolutional_encoder
    enc = get_baseline_convolutional_encoder(16, 8)
    assert enc.precision == 2
    enc.precision = 3
    assert clone_model(enc).precision == 3
    with pytest.raises(ValueError):
        enc.precision = 4


def test_host_stage_casts_in_order_and_recycles_slots(monkeypatch):
    """_HostStage (predict() on numpy batches): every chunk is the float32 cast of its rows, chunks come in order, and
    a slot is only reused after its consumer's copy event -- checked with pinning stubbed out (no CUDA here)."""
    import torch
    from voicemap_b200 import models as M
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
    assert M._cast_plan(256) == [(0, 128), (128, 256)]          # halves (row blocks of a chunk go to all threads)
    assert M._cast_plan(5) == [(0, 5)] and M._cast_plan(40) == [(0, 20), (20, 40)] and M._cast_plan(20) == [(0, 16), (16, 20)]
    for n in (1, 7, 63, 64, 300, 5000):
        plan = M._cast_plan(n)
        assert plan[0][0] == 0 and plan[-1][1] == n and all(a[1] == b[0] for a, b in zip(plan, plan[1:]))
        assert all(0 < hi - lo <= 128 for lo, hi in plan)

    class Event:
        def __init__(self, log, tag):
            self.log, self.tag = log, tag

        def synchronize(self):
            self.log.append(self.tag)

    stage = M._HostStage()
    rng = np.random.default_rng(0)
    batch = rng.normal(size=(1200, 37, 1))  # float64, 10 chunks of 128 rows (the cap) over 4 slots, last one ragged
    waited, seen = [], []
    out = np.zeros((1200, 37), dtype=np.float32)
    for lo, hi, view, slot in stage.chunks(batch[:, :, 0]):
        assert slot[1] is None               # the previous user's event was awaited and cleared before the cast
        out[lo:hi] = view.numpy()
        seen.append((lo, hi))
        slot[1] = Event(waited, lo)
    assert seen == [(lo, min(lo + 128, 1200)) for lo in range(0, 1200, 128)]
    np.testing.assert_array_equal(out, batch[:, :, 0].astype(np.float32))
    assert waited == [lo for lo, _ in seen[:len(seen) - 4]]      # each recycled slot waited for exactly its own copy
    # integer and float32 inputs, second call on the same stage (left-over events are awaited first)
    ints = rng.integers(-5, 5, size=(9, 4))
    got = np.concatenate([view.numpy().copy() for _, _, view, _ in stage.chunks(ints)])
    np.testing.assert_array_equal(got, ints.astype(np.float32))


def test_staging_copy_plain_and_pooled_agree():
    """training._host_copy: the plain assignment and the four-thread row-block form fill the staging buffer alike (any
    source dtype, ragged row counts); which one runs is decided by measurement on the host."""
    from voicemap_b200 import training as T
    rng = np.random.default_rng(0)
    saved = (T._COPY_POOL, T._COPY_THREADS)
    try:
        for rows, cols, dtype in ((128, 12000, np.float32), (37, 20011, np.float64), (9, 70000, np.float32)):
            src = rng.standard_normal((rows, cols)).astype(dtype)
            want = src.astype(np.float32)
            for threads in (1, 4):
                from concurrent.futures import ThreadPoolExecutor
                T._COPY_THREADS = threads
                T._COPY_POOL = ThreadPoolExecutor(max_workers=4) if threads == 4 else None
                dst = np.full((rows, cols), np.nan, dtype=np.float32)
                T._host_copy(dst, src)
                assert np.array_equal(dst, want)
        T._COPY_POOL, T._COPY_THREADS = None, None      # first use: measures, keeps one of the two, result is right
        dst = np.zeros((64, 12000), dtype=np.float32)
        src = rng.standard_normal((64, 12000))
        T._host_copy(dst, src)
        assert np.array_equal(dst, src.astype(np.float32)) and T._COPY_THREADS in (1, 4)
    finally:
        T._COPY_POOL, T._COPY_THREADS = saved
