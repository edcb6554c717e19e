"""Pins the CPU oracle: the reference's one numeric known-answer (whitening), a literal transcription of its
whitening arithmetic, and independent pure-numpy loop restatements of every layer semantic listed in
SURVEY.md 8(c) (SAME-pad asymmetry, ReLU before BN, negative-gamma BN, VALID pool tail drop, label polarity)."""
import math

import numpy as np
import torch

from oracle import voicemap_oracle as O


def test_whiten_matches_literal_transcription():
    rng = np.random.default_rng(0)
    batch = rng.normal(0.3, 2.0, size=(5, 400, 1))
    np.testing.assert_allclose(O.whiten(batch), O.whiten_literal(batch), rtol=1e-12, atol=1e-15)


def test_whiten_known_answer_reference_test():
    # tests/tests.py:71-90 of the reference: the same clip stacked twice -> zero mean, RMS 0.038021 (ddof-free)
    rng = np.random.default_rng(1)
    clip = rng.normal(0.0, 0.2, size=(48000,))
    clip -= clip.mean()
    batch = np.stack([clip, clip])[:, :, np.newaxis]
    w = O.whiten(batch)
    assert abs(w[0].mean()) < 1e-9
    assert math.isclose(np.sqrt(np.power(w[0, :, 0], 2).mean()), 0.038021, rel_tol=1e-6)


def test_whiten_global_scale_quirk():
    # F8: the scale is one scalar for the whole batch, computed on the un-centred batch
    rng = np.random.default_rng(2)
    a = rng.normal(0, 1.0, size=(1, 100, 1))
    b = rng.normal(0, 10.0, size=(1, 100, 1))
    w = O.whiten(np.concatenate([a, b]))
    r = np.sqrt((w ** 2).mean(axis=1))[:, 0]
    assert r[1] / r[0] > 5  # not per-sample RMS normalisation


def test_preprocess_decimates_then_whitens():
    rng = np.random.default_rng(3)
    x = rng.normal(size=(2, 64, 1))
    out = O.preprocess_instances(4)(x)
    assert out.shape == (2, 16, 1)
    np.testing.assert_allclose(out, O.whiten(x[:, ::4, :]))


def _conv_same_relu_loops(x, w, b):
    n, l, cin = x.shape
    k, _, cout = w.shape
    left = (k - 1) // 2
    y = np.zeros((n, l, cout))
    for i in range(n):
        for p in range(l):
            for t in range(k):
                q = p + t - left
                if 0 <= q < l:
                    y[i, p] += x[i, q] @ w[t]
            y[i, p] += b
    return np.maximum(y, 0)


def test_conv_same_padding_k32_is_15_left_16_right():
    rng = np.random.default_rng(4)
    x = rng.normal(size=(1, 40, 1))
    w = rng.normal(size=(32, 1, 3))
    b = rng.normal(size=(3,))
    ref = _conv_same_relu_loops(x, w, b)
    got = O.conv1d_same_relu(torch.tensor(x), torch.tensor(w), torch.tensor(b)).numpy()
    np.testing.assert_allclose(got, ref, rtol=1e-10, atol=1e-12)
    # an impulse at position p lights tap t at output p - t + 15
    x = np.zeros((1, 40, 1)); x[0, 20, 0] = 1.0
    w = np.zeros((32, 1, 1)); w[0, 0, 0] = 1.0
    got = O.conv1d_same_relu(torch.tensor(x), torch.tensor(w), torch.zeros(1, dtype=torch.float64)).numpy()
    assert got[0, 35, 0] == 1.0 and got.sum() == 1.0


def test_conv_k3_matches_loops():
    rng = np.random.default_rng(5)
    x = rng.normal(size=(2, 17, 4))
    w = rng.normal(size=(3, 4, 5))
    b = rng.normal(size=(5,))
    got = O.conv1d_same_relu(torch.tensor(x), torch.tensor(w), torch.tensor(b)).numpy()
    np.testing.assert_allclose(got, _conv_same_relu_loops(x, w, b), rtol=1e-10, atol=1e-12)


def test_bn_eval_negative_gamma_and_relu_order():
    x = torch.tensor([[[-1.0, 2.0]]], dtype=torch.float64)
    t64 = lambda v: torch.tensor(v, dtype=torch.float64)  # noqa: E731
    y = O.batchnorm_eval(x, t64([-2.0, 1.0]), t64([0.5, 0.0]), t64([1.0, 1.0]), t64([3.0, 0.0]), eps=1e-3)
    s0 = -2.0 / math.sqrt(3.0 + 1e-3)
    s1 = 1.0 / math.sqrt(1e-3)
    np.testing.assert_allclose(y.numpy()[0, 0], [(-1 - 1) * s0 + 0.5, (2 - 1) * s1], rtol=1e-12)


def test_bn_train_uses_biased_batch_moments():
    rng = np.random.default_rng(6)
    x = torch.tensor(rng.normal(size=(3, 7, 2)))
    y, m, v = O.batchnorm_train(x, torch.ones(2, dtype=torch.float64), torch.zeros(2, dtype=torch.float64))
    flat = x.numpy().reshape(-1, 2)
    np.testing.assert_allclose(m.numpy(), flat.mean(0))
    np.testing.assert_allclose(v.numpy(), flat.var(0))  # ddof=0
    np.testing.assert_allclose(y.numpy().reshape(-1, 2).mean(0), 0, atol=1e-12)


def test_maxpool_valid_drops_tail():
    x = torch.arange(7, dtype=torch.float64).reshape(1, 7, 1)
    np.testing.assert_array_equal(O.maxpool1d_valid(x, 2).numpy()[0, :, 0], [1, 3, 5])
    np.testing.assert_array_equal(O.maxpool1d_valid(x, 4).numpy()[0, :, 0], [3])


def test_encoder_shapes_and_length_chain():
    p = O.init_encoder_params(16, 8, seed=0)
    x = O.synthetic_clips(2, 6000)  # 1.5 s @ 4 kHz: 6000 -> 1500 -> 750 -> 375 -> 187 (tail dropped)
    emb, inter, gmax, _ = O.encoder_forward(x, p, torch.float32, return_intermediates=True)
    assert [t.shape for t in inter] == [(2, 1500, 16), (2, 750, 32), (2, 375, 48), (2, 187, 64)]
    assert gmax.shape == (2, 64) and emb.shape == (2, 8)
    np.testing.assert_allclose(emb, gmax @ p["dense_kernel"] + p["dense_bias"], rtol=1e-5, atol=1e-6)


def test_encoder_fp32_close_to_fp64():
    p = O.init_encoder_params(32, 16, seed=1, randomize_bn=True, random_bias=True)
    x = O.synthetic_clips(2, 2048)
    e32 = O.encoder_forward(x, p, torch.float32)
    e64 = O.encoder_forward(x, p, torch.float64)
    assert np.abs(e32 - e64).max() / np.abs(e64).max() < 1e-5


def test_glorot_limits_and_param_count():
    shapes = O.encoder_layer_shapes(128, 64)
    trainable = sum(int(np.prod(s)) for k, s in shapes.items() if not (k.endswith("mean") or k.endswith("var")))
    assert trainable == 1_023_808
    rng = np.random.default_rng(0)
    w = O.glorot_uniform((3, 128, 256), rng)
    assert np.abs(w).max() <= math.sqrt(6.0 / (3 * 128 + 3 * 256))


def test_siamese_head_and_losses():
    e1 = np.array([[1.0, 2.0, 2.0], [0.0, 0.0, 0.0]])
    e2 = np.array([[0.0, 0.0, 0.0], [0.0, 0.0, 0.0]])
    p, d = O.siamese_head(e1, e2, w=2.0, b=-1.0)
    np.testing.assert_allclose(d[:, 0], [3.0, 0.0])
    np.testing.assert_allclose(p[:, 0], [1 / (1 + math.exp(-5.0)), 1 / (1 + math.exp(1.0))])
    pw, a = O.siamese_head(e1, e2, w=np.array([1.0, -1.0, 0.5]), b=0.25, distance_metric="weighted_l1")
    np.testing.assert_allclose(pw[:, 0], [1 / (1 + math.exp(-(1 - 2 + 1 + 0.25))), 1 / (1 + math.exp(-0.25))])
    # label polarity: 0 = same speaker -> penalise large p; 1 = different -> penalise p below the margin
    y = np.array([[0.0], [1.0]])
    pred = np.array([[0.2], [0.7]])
    assert math.isclose(O.contrastive_loss(y, pred), (0.2 ** 2 + 0.3 ** 2) / 2)
    assert math.isclose(O.binary_crossentropy(y, pred), (-math.log(0.8) - math.log(0.7)) / 2)
    # keras clips probabilities to [1e-7, 1 - 1e-7]
    assert math.isclose(O.binary_crossentropy(np.array([[1.0]]), np.array([[0.0]])), -math.log(1e-7))


def test_n_shot_predict_argmin_convention():
    q = np.array([[1.0, 0.0]])
    support = np.array([[0.9, 0.1], [1.1, -0.1], [-1.0, 0.0], [-1.0, 0.2]])
    for dist in ("euclidean", "cosine", "dot_product"):
        assert np.argmin(O.n_shot_predict(q, support, n=2, k=2, distance=dist)) == 0


def test_keras_adam_step_matches_closed_form():
    p = {"w": np.array([1.0, -2.0])}
    g = {"w": np.array([3.0, 4.0])}   # norm 5 -> clipped to norm 1
    m = {"w": np.zeros(2)}
    v = {"w": np.zeros(2)}
    O.keras_adam_step(p, g, m, v, t=1)
    gc = np.array([0.6, 0.8])
    lr_t = 1e-3 * math.sqrt(1 - 0.999) / (1 - 0.9)
    expect = np.array([1.0, -2.0]) - lr_t * (0.1 * gc) / (np.sqrt(0.001 * gc ** 2) + 1e-7)
    np.testing.assert_allclose(p["w"], expect, rtol=1e-12)


def test_pool_winner_selection_reproduces_the_plain_gradient():
    """Gradient tests hand the device's pool winners to the oracle (maxpool1d_valid ``select``, GlobalMaxPool
    ``gmax_select``).  With the oracle's OWN winners the selected forward/backward must equal the plain one, the slack
    must be zero, and a selection that is not a maximum must show up in the slack."""
    params = O.init_encoder_params(16, 8, seed=0, random_bias=True)
    rng = np.random.default_rng(3)
    for i in range(1, 5):
        params[f"bn{i}_gamma"] = rng.uniform(-1.2, 1.5, params[f"bn{i}_gamma"].shape).astype(np.float32)
    x1, x2 = O.synthetic_clips(2, 512, seed=1), O.synthetic_clips(2, 512, seed=2)
    y = np.array([0.0, 1.0])
    plain = O.siamese_train_step_grads(params, np.array([0.05]), np.array([-0.3]), x1, x2, y)
    selects = []
    for branch in range(2):
        flags = []
        for b, (u, pool) in enumerate(zip(plain["u"][branch], O.POOLS)):
            lo = u.shape[1] // pool
            win = u[:, :lo * pool].reshape(u.shape[0], lo, pool, -1)
            pick = np.where(params[f"bn{b + 1}_gamma"][None, None, :] < 0, win.argmin(2), win.argmax(2))
            f = np.zeros_like(win, dtype=bool)
            np.put_along_axis(f, pick[:, :, None, :], True, 2)
            full = np.zeros(u.shape, bool)
            full[:, :lo * pool] = f.reshape(u.shape[0], lo * pool, -1)
            flags.append(full)
        selects.append(flags)
    chosen = O.siamese_train_step_grads(params, np.array([0.05]), np.array([-0.3]), x1, x2, y, pool_selects=selects)
    assert chosen["select_slack"] == 0.0 and abs(chosen["loss"] - plain["loss"]) < 1e-14
    for k in plain["grads"]:
        np.testing.assert_allclose(chosen["grads"][k], plain["grads"][k], rtol=0, atol=1e-13)
    wrong = [[f.copy() for f in side] for side in selects]
    wrong[0][0][:, :4, :] = np.roll(wrong[0][0][:, :4, :], 1, axis=1)       # first window of block 1: another element
    moved = O.siamese_train_step_grads(params, np.array([0.05]), np.array([-0.3]), x1, x2, y, pool_selects=wrong)
    assert moved["select_slack"] > 1e-3
