"""Golden vectors produced by EXECUTING the reference's own voicemap/models.py and voicemap/utils.py (against a numpy
stand-in for Keras: tests/golden/keras_standin.py, generator tests/golden/make_reference_golden.py).  They pin the
architecture, the siamese wiring and heads, the contrastive loss, preprocessing and the n-shot decision rules to the
reference's code; the arithmetic inside each Keras layer is restated in the stand-in, independently of the oracle
(numpy vs torch, different conv / pooling formulations).

CPU: the oracle and the host-side utilities against the fixture.  GPU (-m gpu): the CUDA path against the fixture.
"""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import reference_cases as RC  # noqa: E402
from oracle import voicemap_oracle as O  # noqa: E402
from voicemap_b200 import utils  # noqa: E402


@pytest.fixture(scope="module")
def golden():
    return np.load(RC.FIXTURE)


def _rel(a, b):
    return float(np.linalg.norm(np.asarray(a, np.float64) - b) / np.linalg.norm(b))


def _pairs(name, golden):
    filters, emb, n, length, _, _ = RC.ENCODER_CASES[name]
    params, x = RC.encoder_inputs(name)
    half = n // 2
    y = (np.arange(half) % 2).astype(np.float64)[:, None]
    return params, x, half, y, emb, float(golden[f"{name}_head_scale"])


# ------------------------------------------------------------------------------------------------ structure
def test_reference_layer_sequence(golden):
    """What voicemap/models.py:6-41 actually builds, recorded while it ran."""
    block = ["Conv1D", "BatchNormalization", "SpatialDropout1D", "MaxPool1D"]
    for name in RC.ENCODER_CASES:
        assert golden[f"{name}_layer_kinds"].tolist() == block * 4 + ["GlobalMaxPool1D", "Dense"]


# ------------------------------------------------------------------------------------------------ oracle vs reference
@pytest.mark.parametrize("name", list(RC.ENCODER_CASES))
def test_oracle_encoder_equals_reference_executed(golden, name):
    params, x = RC.encoder_inputs(name)
    emb64, inter, _, _ = O.encoder_forward(x, params, torch.float64, return_intermediates=True)
    assert _rel(emb64, golden[f"{name}_emb"]) < 1e-11
    for i, act in enumerate(inter, start=1):
        want = golden[f"{name}_block{i}_head"]
        assert act[:, :24, :].shape == want.shape                     # pooled lengths: tails dropped identically
        assert _rel(act[:, :24, :], want) < 1e-11
    emb32 = O.encoder_forward(x, params, torch.float32)
    per_clip = np.linalg.norm(emb32 - golden[f"{name}_emb"], axis=1) / np.linalg.norm(golden[f"{name}_emb"], axis=1)
    assert per_clip.max() < 2e-5                                       # the fp32 parity reference itself


@pytest.mark.parametrize("name", list(RC.ENCODER_CASES))
@pytest.mark.parametrize("metric", ["uniform_euclidean", "weighted_l1"])
def test_oracle_siamese_heads_and_contrastive_loss(golden, name, metric):
    params, x, half, y, emb, scale = _pairs(name, golden)
    kernel, bias = RC.head_weights(metric, emb, scale)
    e = O.encoder_forward(x, params, torch.float64)
    prob, _ = O.siamese_head(e[:half], e[half:], kernel, bias, metric)
    want = golden[f"{name}_{metric}_prob"]
    assert prob.shape == want.shape == (half, 1)
    assert 0.02 < want.min() and want.max() < 0.98                     # the comparison is not about a saturated sigmoid
    np.testing.assert_allclose(prob, want, rtol=1e-10)
    np.testing.assert_allclose(O.contrastive_loss(y, prob), golden[f"{name}_{metric}_contrastive"], rtol=1e-10)
    np.testing.assert_allclose(utils.contrastive_loss(y, want), golden[f"{name}_{metric}_contrastive"], rtol=1e-10)


def test_oracle_classifier_softmax(golden):
    params, x = RC.encoder_inputs("f16")
    kernel, bias = RC.classifier_head(RC.ENCODER_CASES["f16"][1])
    logits = O.encoder_forward(x, params, torch.float64) @ kernel + bias
    z = np.exp(logits - logits.max(axis=1, keepdims=True))
    np.testing.assert_allclose(z / z.sum(axis=1, keepdims=True), golden["f16_classifier_prob"], rtol=1e-9, atol=1e-300)


# ------------------------------------------------------------------------------------------------ host utilities
def test_preprocessing_equals_reference_executed(golden):
    raw = RC.raw_clips()
    for impl in (utils, O):
        np.testing.assert_allclose(impl.preprocess_instances(4)(raw), golden["preprocess_ds4"], rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(impl.preprocess_instances(1)(raw), golden["preprocess_ds1"], rtol=1e-12, atol=1e-15)
        np.testing.assert_array_equal(impl.preprocess_instances(4, whitening=False)(raw),
                                      golden["preprocess_ds4_no_whiten"])
        np.testing.assert_allclose(impl.whiten(raw), golden["whiten"], rtol=1e-12, atol=1e-15)
    pre = utils.BatchPreProcessor("siamese", utils.preprocess_instances(4))
    (left, right), labels = pre(([raw, raw[::-1]], np.array([[0.0], [1.0], [1.0]])))
    np.testing.assert_allclose(left, golden["batch_preprocessor_left"], rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(right, golden["batch_preprocessor_right"], rtol=1e-12, atol=1e-15)
    np.testing.assert_array_equal(labels, golden["batch_preprocessor_labels"])


class _OracleEncoder:
    """Duck-typed Keras Sequential backed by the fp64 oracle (CPU stand-in for EncoderModel in the n-shot tests)."""

    def __init__(self, params, head=None):
        self.params, self.head = params, head

    def predict(self, x):
        e = O.encoder_forward(np.asarray(x, dtype=np.float64), self.params, torch.float64)
        if self.head is not None:
            logits = e @ self.head[0] + self.head[1]
            z = np.exp(logits - logits.max(axis=1, keepdims=True))
            return z / z.sum(axis=1, keepdims=True)
        return e

    # the slice of the Keras API voicemap/utils.py:143-145 uses on a classifier
    def _clone(self):
        return _OracleEncoder(None, self.head)

    def get_weights(self):
        return [self.params, self.head]

    def set_weights(self, weights):
        self.params, self.head = weights

    def pop(self):
        self.head = None


class _OracleSiamese:
    def __init__(self, params, kernel, bias):
        self.encoder = _OracleEncoder(params)
        self.layers = [None, None, self.encoder, None, None, None]
        self.kernel, self.bias = kernel, bias

    def predict(self, xs):
        e1, e2 = self.encoder.predict(xs[0]), self.encoder.predict(xs[1])
        return O.siamese_head(e1, e2, self.kernel, self.bias)[0]


def _nshot_model(network_type, params, emb, scale, make_siamese, make_classifier):
    if network_type == "siamese":
        return make_siamese(params, *RC.head_weights("uniform_euclidean", emb, scale))
    return make_classifier(params, RC.classifier_head(emb))


@pytest.mark.parametrize("case", range(len(RC.NSHOT_CASES)))
def test_n_shot_evaluation_equals_reference_executed(golden, case):
    """voicemap_b200.utils.n_shot_task_evaluation vs the reference's function on the same task stream and the same
    network (here evaluated by the oracle): identical number of solved tasks, for all five decision paths."""
    network_type, n, k, distance, tasks, seed, noise = RC.NSHOT_CASES[case]
    assert bool(golden["nshot_robust"][case])
    params, _ = RC.encoder_inputs("f16")
    emb = RC.ENCODER_CASES["f16"][1]
    model = _nshot_model(network_type, params, emb, float(golden["nshot_head_scale"]), _OracleSiamese,
                         lambda p, head: _OracleEncoder(p, head))
    pre = utils.BatchPreProcessor("siamese", utils.preprocess_instances(RC.NSHOT_DOWNSAMPLING))
    got = utils.n_shot_task_evaluation(model, RC.TaskDataset(seed, noise=noise), pre, tasks, n, k,
                                       network_type=network_type, distance=distance)
    assert got == int(golden["nshot_correct"][case])


# ------------------------------------------------------------------------------------------------ CUDA path
def _gpu_encoder(name, precision=None):
    from voicemap_b200.models import get_baseline_convolutional_encoder
    filters, emb, _, _, _, _ = RC.ENCODER_CASES[name]
    params, x = RC.encoder_inputs(name)
    encoder = get_baseline_convolutional_encoder(filters, emb, dropout=0.05)
    encoder.set_weights(RC.keras_weight_list(params))                 # Keras get_weights() order, as the reference
    if precision is not None:
        encoder.precision = precision
    return encoder, params, x


@pytest.mark.gpu
@pytest.mark.parametrize("precision", [2, 3])
@pytest.mark.parametrize("name", list(RC.ENCODER_CASES))
def test_gpu_encoder_equals_reference_executed(golden, name, precision):
    encoder, _, x = _gpu_encoder(name, precision)
    got = encoder.predict(x)
    want = golden[f"{name}_emb"]
    per_clip = np.linalg.norm(got - want, axis=1) / np.linalg.norm(want, axis=1)
    assert per_clip.max() < 1e-4, per_clip                            # north-star tolerance: 1e-4 relative
    assert np.abs(got - want).max() / np.abs(want).max() < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("metric", ["uniform_euclidean", "weighted_l1"])
def test_gpu_siamese_equals_reference_executed(golden, metric):
    from voicemap_b200.models import build_siamese_net
    name = "f128"
    encoder, _, x = _gpu_encoder(name)
    _, _, half, y, emb, scale = _pairs(name, golden)
    siamese = build_siamese_net(encoder, (x.shape[1], 1), distance_metric=metric)
    assert siamese.layers[2] is encoder
    kernel, bias = RC.head_weights(metric, emb, scale)
    siamese.set_weights(encoder.get_weights() + [kernel.astype(np.float32), bias.astype(np.float32)])
    prob = siamese.predict([x[:half], x[half:]])
    want = golden[f"{name}_{metric}_prob"]
    assert prob.shape == want.shape
    # probabilities and losses within 1e-4 RELATIVE of the reference, end to end through the CUDA
    # encoder (siamese eval runs the fp16 x 3 arithmetic: the head works on differences of embeddings)
    assert np.abs(prob - want).max() <= 1e-4 * np.abs(want).max()
    np.testing.assert_allclose(prob, want, rtol=1e-4)
    # the losses through the fused head + loss kernel (vm_pair_head_loss_fwd), not a host formula
    from voicemap_b200.keras_compat import Adam
    siamese.compile(loss=utils.contrastive_loss, optimizer=Adam())
    got = siamese.test_on_batch([x[:half], x[half:]], y)
    ref_loss = float(golden[f"{name}_{metric}_contrastive"])
    assert abs(got - ref_loss) <= 1e-4 * abs(ref_loss), (got, ref_loss)
    siamese.compile(loss="binary_crossentropy", optimizer=Adam(), metrics=["accuracy"])
    got_bce, got_acc = siamese.test_on_batch([x[:half], x[half:]], y)
    w64 = want.astype(np.float64)
    pc = np.clip(w64, 1e-7, 1 - 1e-7)                       # keras binary_crossentropy on the reference's probabilities
    ref_bce = float(np.mean(-y * np.log(pc) - (1 - y) * np.log(1 - pc)))
    assert abs(got_bce - ref_bce) <= 1e-4 * abs(ref_bce), (got_bce, ref_bce)
    assert got_acc == float(np.mean((w64 > 0.5) == (y > 0.5)))


@pytest.mark.gpu
@pytest.mark.parametrize("case", range(len(RC.NSHOT_CASES)))
def test_gpu_n_shot_evaluation_equals_reference_executed(golden, case):
    from voicemap_b200.keras_compat import Dense
    from voicemap_b200.models import build_siamese_net, get_baseline_convolutional_encoder
    network_type, n, k, distance, tasks, seed, noise = RC.NSHOT_CASES[case]
    filters, emb = RC.ENCODER_CASES["f16"][:2]
    params, _ = RC.encoder_inputs("f16")
    samples = RC.TaskDataset(0).samples // RC.NSHOT_DOWNSAMPLING
    if network_type == "siamese":
        encoder = get_baseline_convolutional_encoder(filters, emb)
        encoder.set_weights(RC.keras_weight_list(params))
        model = build_siamese_net(encoder, (samples, 1))
        kernel, bias = RC.head_weights("uniform_euclidean", emb, float(golden["nshot_head_scale"]))
        model.set_weights(encoder.get_weights() + [kernel.astype(np.float32), bias.astype(np.float32)])
    else:
        model = get_baseline_convolutional_encoder(filters, emb, (samples, 1))
        model.add(Dense(7, activation="softmax"))
        kernel, bias = RC.classifier_head(emb)
        model.set_weights(RC.keras_weight_list(params) + [kernel.astype(np.float32), bias.astype(np.float32)])
    pre = utils.BatchPreProcessor("siamese", utils.preprocess_instances(RC.NSHOT_DOWNSAMPLING))
    for per_launch in (1, 8):
        got = utils.n_shot_task_evaluation(model, RC.TaskDataset(seed, noise=noise), pre, tasks, n, k,
                                           network_type=network_type, distance=distance, tasks_per_launch=per_launch)
        assert got == int(golden["nshot_correct"][case]), (per_launch, got)
