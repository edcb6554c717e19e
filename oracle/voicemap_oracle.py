"""CPU oracle for the voicemap 1D-conv speaker-embedding hot path.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this
module.  The product path (``voicemap_b200``) never does: it fails loudly when
the CUDA library is missing.

PARITY UNPINNED against Keras itself.  The reference (oscarknagg/voicemap @ dd79c69) is Python 2.7 on
Keras 2.2.2 / tensorflow-gpu 1.10.1, neither of which exists in this image, and its own tests
(``tests/tests.py``) never build a model, so no golden activation computed BY KERAS exists.
What is pinned:
  * against the reference's own source, executed in the build container: ``voicemap/models.py``
    (imported unchanged) and ``voicemap/utils.py`` run against a numpy stand-in for the Keras symbols
    they use (``tests/golden/keras_standin.py``); the resulting embeddings, block activations, siamese
    outputs of both heads, contrastive losses, preprocessing outputs and n-shot decision counts are
    committed as ``tests/golden/reference_executed.npz`` and this oracle reproduces them to 1e-11
    (``tests/test_reference_golden.py``).  That fixes architecture, wiring, heads, loss,
    preprocessing and evaluation rules to the reference's code; the arithmetic inside each Keras
    layer is a restatement there too (independent of this file: numpy, other formulations);
  * ``whiten`` against the one known-answer the reference's tests hold
    (``tests/tests.py:71-90``: zero mean, RMS 0.038021 for a repeated clip) and
    against a literal transcription of the reference's tile/transpose arithmetic;
  * layer semantics against independent pure-numpy loops (``tests/test_oracle.py``).

Every function cites the reference file:line it restates.  Arithmetic that the
reference delegates to Keras/TF (conv, BN, pooling, losses, Adam) follows the
Keras 2.2.2 semantics listed in SURVEY.md section 8(c).

Layouts are Keras': activations (N, L, C) channels-last, conv kernel
(K, Cin, Cout), dense kernel (in, out).  ``dtype`` selects fp32 (the parity
reference) or fp64 (the adjudicator).
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-3          # keras.layers.BatchNormalization default epsilon
BN_MOMENTUM = 0.99     # keras default momentum
WHITEN_RMS = 0.038021  # voicemap/utils.py:88


# ----------------------------------------------------------------------------
# parameters
# ----------------------------------------------------------------------------
def encoder_layer_shapes(filters: int, embedding_dimension: int):
    """Shapes of every weight of get_baseline_convolutional_encoder
    (voicemap/models.py:6-41) in Keras order."""
    f = filters
    shapes = OrderedDict()
    cin = 1
    for i, (k, mult) in enumerate(((32, 1), (3, 2), (3, 3), (3, 4)), start=1):
        cout = mult * f
        shapes[f"conv{i}_kernel"] = (k, cin, cout)
        shapes[f"conv{i}_bias"] = (cout,)
        shapes[f"bn{i}_gamma"] = (cout,)
        shapes[f"bn{i}_beta"] = (cout,)
        shapes[f"bn{i}_mean"] = (cout,)
        shapes[f"bn{i}_var"] = (cout,)
        cin = cout
    shapes["dense_kernel"] = (4 * f, embedding_dimension)
    shapes["dense_bias"] = (embedding_dimension,)
    return shapes


def glorot_uniform(shape, rng: np.random.Generator):
    """keras.initializers.glorot_uniform: U(-l, l), l = sqrt(6 / (fan_in + fan_out)).
    Conv kernel (K, Cin, Cout): fan_in = K*Cin, fan_out = K*Cout."""
    if len(shape) == 2:
        fan_in, fan_out = shape
    else:
        receptive = int(np.prod(shape[:-2]))
        fan_in, fan_out = shape[-2] * receptive, shape[-1] * receptive
    limit = math.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-limit, limit, size=shape).astype(np.float32)


def init_encoder_params(filters, embedding_dimension, seed=0, randomize_bn=False,
                        random_bias=False):
    """Keras-default initialisation (glorot-uniform kernels, zero biases, gamma=1,
    beta=0, moving mean 0, moving var 1).  ``randomize_bn`` applies the parity
    stress distribution of SURVEY.md 8(d): gamma~U(-1.5,1.5) (incl. negatives),
    beta~N(0,.3), mean~N(0,.5), var~logU(1e-6,30)."""
    rng = np.random.default_rng(seed)
    p = OrderedDict()
    for name, shape in encoder_layer_shapes(filters, embedding_dimension).items():
        if name.endswith("kernel"):
            p[name] = glorot_uniform(shape, rng)
        elif name.endswith("bias"):
            p[name] = (rng.normal(0, 0.05, shape).astype(np.float32) if random_bias
                       else np.zeros(shape, np.float32))
        elif name.endswith("gamma"):
            p[name] = (rng.uniform(-1.5, 1.5, shape).astype(np.float32) if randomize_bn
                       else np.ones(shape, np.float32))
        elif name.endswith("beta"):
            p[name] = (rng.normal(0, 0.3, shape).astype(np.float32) if randomize_bn
                       else np.zeros(shape, np.float32))
        elif name.endswith("mean"):
            p[name] = (rng.normal(0, 0.5, shape).astype(np.float32) if randomize_bn
                       else np.zeros(shape, np.float32))
        elif name.endswith("var"):
            p[name] = (np.exp(rng.uniform(math.log(1e-6), math.log(30.0), shape)).astype(np.float32)
                       if randomize_bn else np.ones(shape, np.float32))
    return p


def synthetic_clips(n, length, seed=1234, padded=False):
    """SURVEY.md 8(d) synthetic input: 0.038021 * randn(N, L, 1) fp32; variant B
    zeroes a random prefix+suffix per clip like pad=True stochastic padding
    (voicemap/librispeech.py:114-124)."""
    g = torch.Generator().manual_seed(seed)
    x = WHITEN_RMS * torch.randn(n, length, 1, generator=g, dtype=torch.float32)
    if padded:
        rng = np.random.default_rng(seed + 1)
        for i in range(n):
            less = int(rng.integers(0, max(length // 3, 1)))
            before = int(rng.integers(0, less + 1))
            x[i, :before] = 0
            if less - before:
                x[i, length - (less - before):] = 0
    return x.numpy()


# ----------------------------------------------------------------------------
# preprocessing (numpy, as the reference)
# ----------------------------------------------------------------------------
def whiten(batch, rms=WHITEN_RMS):
    """voicemap/utils.py:88-101.  Per-sample mean removal, then ONE global scalar
    rms / sqrt(mean(batch**2)) taken over the whole un-centred batch (:98)."""
    if batch.ndim != 3:
        raise ValueError("Input must be a 3D array of shape (n_segments, n_timesteps, 1).")
    sample_wise_mean = batch.mean(axis=1, keepdims=True)           # :94-95
    rescale = rms / np.sqrt(np.power(batch, 2).mean())             # :98
    return (batch - sample_wise_mean) * rescale                     # :95,:99


def whiten_literal(batch, rms=WHITEN_RMS):
    """Line-by-line transcription of voicemap/utils.py:93-101 (tile/transpose
    form) used to pin ``whiten`` above."""
    sample_wise_mean = batch.mean(axis=1)
    whitened = batch - np.tile(sample_wise_mean, (1, 1, batch.shape[1])).transpose((1, 2, 0))
    rescaling = rms / np.sqrt(np.power(batch, 2).mean())
    whitened = whitened * np.tile(rescaling, (1, 1, batch.shape[1])).transpose((1, 2, 0))
    return whitened


def preprocess_instances(downsampling, whitening=True):
    """voicemap/utils.py:22-34: instances[:, ::downsampling, :] then whiten."""
    def _pre(instances):
        instances = instances[:, ::downsampling, :]
        if whitening:
            instances = whiten(instances)
        return instances
    return _pre


# ----------------------------------------------------------------------------
# layers (torch CPU)
# ----------------------------------------------------------------------------
def _t(a, dtype):
    return torch.as_tensor(np.asarray(a)).to(dtype)


def conv1d_same_relu(x, kernel, bias, relu_mask=None):
    """keras.layers.Conv1D(filters, K, padding='same', activation='relu')
    (voicemap/models.py:13,16,22,27,32).  Cross-correlation, stride 1, zero pad
    left (K-1)//2, right K-1-left (K=32 -> 15/16).
    ``relu_mask`` (N, L, Cout), optional: use this 0/1 pattern instead of (pre-activation > 0).  Gradient tests pass
    the device's own activation pattern so that a pre-activation within rounding of zero (where fp32 and fp64
    legitimately disagree on the ReLU branch) does not turn into a spurious gradient mismatch."""
    k = kernel.shape[0]
    left = (k - 1) // 2
    right = k - 1 - left
    xin = F.pad(x.transpose(1, 2), (left, right))                  # (N, Cin, L+K-1)
    w = kernel.permute(2, 1, 0).contiguous()                       # (Cout, Cin, K)
    y = F.conv1d(xin, w, bias)
    if relu_mask is not None:
        return (y * relu_mask.transpose(1, 2)).transpose(1, 2)
    return torch.relu(y).transpose(1, 2)                           # (N, L, Cout)


def batchnorm_eval(x, gamma, beta, mean, var, eps=BN_EPS):
    """keras BatchNormalization inference on a 3-D input (non-fused TF path):
    x*(gamma*rsqrt(var+eps)) + (beta - mean*gamma*rsqrt(var+eps))
    (voicemap/models.py:17,23,28,33)."""
    s = gamma * torch.rsqrt(var + eps)
    return x * s + (beta - mean * s)


def batchnorm_train(x, gamma, beta, eps=BN_EPS):
    """Training-mode BN: biased batch moments over (N, L) per channel
    (tf.nn.moments) then tf.nn.batch_normalization.  Returns y, mean, var."""
    mean = x.mean(dim=(0, 1))
    var = ((x - mean) ** 2).mean(dim=(0, 1))
    s = gamma * torch.rsqrt(var + eps)
    return x * s + (beta - mean * s), mean, var


def maxpool1d_valid(x, pool, select=None, slack=None):
    """keras MaxPool1D(pool, pool), padding 'valid': L_out = floor(L/pool), the
    tail is dropped (voicemap/models.py:19,25,30,35).
    ``select`` (N, L, C) bool, optional: take the flagged element of every window (exactly one per window) instead of
    the maximum.  Gradient tests pass the device's own arg-max pattern: where two values of a window agree to within
    rounding, fp32 and fp64 legitimately pick different winners, and the gradient of the pool -- which lands on the
    winner only -- would differ by a whole entry without either side being wrong.  ``slack`` (a list) receives
    max((window maximum - selected value) / max|x|), so that a caller can check that every selection IS a maximum up
    to rounding."""
    n, l, c = x.shape
    lo = l // pool
    win = x[:, :lo * pool, :].reshape(n, lo, pool, c)
    if select is None:
        return win.amax(dim=2)
    sel = select[:, :lo * pool, :].reshape(n, lo, pool, c)
    if not bool((sel.sum(dim=2) == 1).all()):
        raise ValueError("maxpool1d_valid: `select` must flag exactly one element per window")
    picked = (win * sel.to(win.dtype)).sum(dim=2)
    if slack is not None:
        slack.append(float(((win.amax(dim=2) - picked).max() / x.abs().max()).item()))
    return picked


POOLS = (4, 2, 2, 2)  # voicemap/models.py:19,25,30,35


def encoder_forward(x, params, dtype=torch.float32, training=False, return_intermediates=False,
                    pools=POOLS):
    """get_baseline_convolutional_encoder forward (voicemap/models.py:6-41):
    4 x [Conv1D+ReLU -> BN -> (SpatialDropout1D: identity in eval / rate 0) -> MaxPool]
    -> GlobalMaxPool1D -> Dense(embedding_dimension) (linear, bias).
    x: (N, L, 1).  Returns (N, emb) numpy (and the per-block pooled outputs)."""
    h = _t(x, dtype)
    inter = []
    stats = []
    for i in range(1, 5):
        h = conv1d_same_relu(h, _t(params[f"conv{i}_kernel"], dtype), _t(params[f"conv{i}_bias"], dtype))
        g, b = _t(params[f"bn{i}_gamma"], dtype), _t(params[f"bn{i}_beta"], dtype)
        if training:
            h, m, v = batchnorm_train(h, g, b)
            stats.append((m.numpy(), v.numpy()))
        else:
            h = batchnorm_eval(h, g, b, _t(params[f"bn{i}_mean"], dtype), _t(params[f"bn{i}_var"], dtype))
        h = maxpool1d_valid(h, pools[i - 1])
        inter.append(h)
    gmax = h.amax(dim=1)                                            # GlobalMaxPool1D  models.py:37
    emb = gmax @ _t(params["dense_kernel"], dtype) + _t(params["dense_bias"], dtype)   # models.py:39
    if return_intermediates:
        return emb.numpy(), [t.numpy() for t in inter], gmax.numpy(), stats
    return emb.numpy()


# ----------------------------------------------------------------------------
# siamese head + losses (numpy)
# ----------------------------------------------------------------------------
def _sigmoid(z):
    return 1.0 / (1.0 + np.exp(-z))


def siamese_head(e1, e2, w, b, distance_metric="uniform_euclidean"):
    """build_siamese_net head (voicemap/models.py:55-69).
    uniform_euclidean: d = sqrt(sum((e1-e2)^2, -1, keepdims)); p = sigmoid(w*d + b), w,b scalar
                       (K.sqrt clips its argument to >= 0).
    weighted_l1:       p = sigmoid(|e1-e2| @ w + b), w (emb, 1)."""
    diff = e1 - e2
    if distance_metric == "uniform_euclidean":
        d = np.sqrt(np.maximum(np.sum(np.square(diff), axis=-1, keepdims=True), 0.0))
        z = d * np.asarray(w).reshape(1, 1) + np.asarray(b).reshape(1, 1)
        return _sigmoid(z), d
    if distance_metric == "weighted_l1":
        a = np.abs(diff)
        z = a @ np.asarray(w).reshape(-1, 1) + np.asarray(b).reshape(1, 1)
        return _sigmoid(z), a
    raise NotImplementedError(distance_metric)


def contrastive_loss(y_true, y_pred, margin=1.0):
    """voicemap/utils.py:77-85 (Hadsell'06): mean((1-y)*p^2 + y*max(margin-p,0)^2);
    y: 0 = same speaker, 1 = different (voicemap/librispeech.py:194)."""
    y_true = np.asarray(y_true, dtype=y_pred.dtype)
    return np.mean((1 - y_true) * np.square(y_pred) + y_true * np.square(np.maximum(margin - y_pred, 0)))


def binary_crossentropy(y_true, y_pred, eps=1e-7):
    """keras 'binary_crossentropy' on probabilities (experiments/train_siamese.py:57):
    clip p to [eps, 1-eps], mean(-y*log p - (1-y)*log(1-p))."""
    y_true = np.asarray(y_true, dtype=y_pred.dtype)
    p = np.clip(y_pred, eps, 1 - eps)
    return np.mean(-y_true * np.log(p) - (1 - y_true) * np.log(1 - p))


def categorical_crossentropy_from_logits(y_onehot, logits):
    """Dense(num_classes, softmax) + 'categorical_crossentropy'
    (experiments/train_classifier.py:112,115)."""
    z = logits - logits.max(axis=-1, keepdims=True)
    logp = z - np.log(np.exp(z).sum(axis=-1, keepdims=True))
    return np.mean(-(y_onehot * logp).sum(axis=-1))


def siamese_forward(x1, x2, params, head_w, head_b, dtype=torch.float32,
                    distance_metric="uniform_euclidean"):
    """build_siamese_net forward (voicemap/models.py:49-79): shared encoder on both inputs."""
    e1 = encoder_forward(x1, params, dtype)
    e2 = encoder_forward(x2, params, dtype)
    p, d = siamese_head(e1, e2, head_w, head_b, distance_metric)
    return p, d, e1, e2


# ----------------------------------------------------------------------------
# n-shot evaluation on embeddings (voicemap/utils.py:156-212)
# ----------------------------------------------------------------------------
def n_shot_predict(query_embedding, support_embeddings, n, k, distance="euclidean"):
    """Distances from the query to the k class-mean support embeddings; the
    reference scores a task correct when argmin == 0."""
    if distance == "euclidean":
        means = support_embeddings.reshape(k, n, -1).mean(axis=1)
        return np.sqrt(np.power(query_embedding.reshape(1, -1) - means, 2).sum(axis=1))
    mags = np.linalg.norm(support_embeddings, axis=1, keepdims=True)
    units = support_embeddings / mags
    mean_units = units.reshape(k, n, -1).mean(axis=1)
    if distance == "cosine":
        q = query_embedding.reshape(-1)
        return 1.0 - (mean_units @ q) / (np.linalg.norm(mean_units, axis=1) * np.linalg.norm(q))
    if distance == "dot_product":
        mean_mags = mags.reshape(k, n).sum(axis=1, keepdims=True) / n
        return -(query_embedding.reshape(1, -1) @ (mean_mags * mean_units).T).reshape(-1)
    raise ValueError("Distance must be in (euclidean, cosine, dot_product)")


# ----------------------------------------------------------------------------
# training-step restatement (autograd) -- keras Adam(clipnorm=1.) semantics
# ----------------------------------------------------------------------------
def keras_adam_step(params, grads, m, v, t, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-7,
                    clipnorm=1.0, decay=0.0):
    """keras.optimizers.Adam.get_updates (Keras 2.2.2) with clipnorm:
    global-norm clip (clip_norm over all grads), lr_t = lr*sqrt(1-b2^t)/(1-b1^t),
    p -= lr_t * m / (sqrt(v) + eps).  Operates in-place on float64/32 numpy dicts."""
    if clipnorm and clipnorm > 0:
        norm = math.sqrt(sum(float(np.sum(np.square(g.astype(np.float64)))) for g in grads.values()))
        if norm > clipnorm:
            grads = {k: g * (clipnorm / norm) for k, g in grads.items()}
    lr0 = lr * (1.0 / (1.0 + decay * (t - 1))) if decay > 0 else lr
    lr_t = lr0 * math.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t)
    for k in params:
        g = grads[k]
        m[k] = beta1 * m[k] + (1 - beta1) * g
        v[k] = beta2 * v[k] + (1 - beta2) * np.square(g)
        params[k] = params[k] - lr_t * m[k] / (np.sqrt(v[k]) + eps)
    return params, m, v


# ----------------------------------------------------------------------------
# training-mode restatement with autograd (the checker for the backward kernels)
# ----------------------------------------------------------------------------
def bn_moving_update(moving_mean, moving_var, mean, var, n, momentum=BN_MOMENTUM, eps=BN_EPS):
    """keras.layers.BatchNormalization.call (Keras 2.2.2, training branch): the batch variance is turned into the
    sample variance var * n / (n - (1 + eps)) before the moving-average update; moving <- moving * momentum +
    stat * (1 - momentum).  (TF's zero_debias bookkeeping of K.moving_average_update is not restated: its extra
    variables are not part of the Keras weights and cannot be pinned from the reference.)"""
    var_unbiased = var * (n / (n - (1.0 + eps)))
    return (moving_mean * momentum + mean * (1.0 - momentum),
            moving_var * momentum + var_unbiased * (1.0 - momentum))


def _encoder_forward_train_torch(x, P, pools=POOLS, dropout_masks=None, keep=None, relu_masks=None, pool_selects=None,
                                 gmax_select=None, slack=None):
    """Train-mode encoder on torch tensors (autograd-capable).  x (N, L, 1); P: dict name -> tensor.
    dropout_masks: optional list of 4 keep-masks (N, 1, C) already scaled by 1/(1-p) (SpatialDropout1D,
    voicemap/models.py:18,24,29,34).  pool_selects: optional list of 4 bool arrays (N, L_b, C_b), the winner of every
    MaxPool window (see maxpool1d_valid); gmax_select (N, C4) int, the winning window of GlobalMaxPool1D.  Returns emb
    and per-block (mean, var, count)."""
    h = x
    stats = []
    for i in range(1, 5):
        h = conv1d_same_relu(h, P[f"conv{i}_kernel"], P[f"conv{i}_bias"], relu_mask=None if relu_masks is None
                             else relu_masks[i - 1])
        if keep is not None:          # expose d loss / d u (post-ReLU, pre-BN) to the kernel-level tests
            h.retain_grad()
            keep.append(h)
        n_red = h.shape[0] * h.shape[1]
        h, m, v = batchnorm_train(h, P[f"bn{i}_gamma"], P[f"bn{i}_beta"])
        stats.append((m.detach(), v.detach(), n_red))
        if dropout_masks is not None and dropout_masks[i - 1] is not None:
            h = h * dropout_masks[i - 1]
        h = maxpool1d_valid(h, pools[i - 1], None if pool_selects is None else pool_selects[i - 1], slack)
    if gmax_select is None:
        gmax = h.amax(dim=1)
    else:
        gmax = torch.gather(h, 1, gmax_select.to(torch.int64)[:, None, :])[:, 0, :]
        if slack is not None:
            slack.append(float(((h.amax(dim=1) - gmax).max() / h.abs().max()).item()))
    return gmax @ P["dense_kernel"] + P["dense_bias"], stats


TRAINABLE_SUFFIXES = ("kernel", "bias", "gamma", "beta")


def siamese_train_step_grads(params, head_w, head_b, x1, x2, y, loss="binary_crossentropy",
                             distance_metric="uniform_euclidean", dtype=torch.float64, dropout_masks=(None, None),
                             relu_masks=(None, None), pool_selects=(None, None), gmax_selects=(None, None)):
    """One training-mode forward/backward of build_siamese_net (voicemap/models.py:49-79) with the loss of
    experiments/train_siamese.py:57 ('binary_crossentropy') or siamese_contrastive_loss.py:70 (contrastive_loss).
    The shared encoder is applied once per branch, so BN batch statistics are per branch.
    pool_selects / gmax_selects: per branch, the device's own pool winners (see maxpool1d_valid); ``select_slack`` in
    the result is how far any of them is from the true maximum, relative to the tensor's largest entry.
    Returns dict(loss, prob, grads {name: array}, head grads, stats [branch][block] -> (mean, var, n))."""
    P = {k: _t(v, dtype).clone().requires_grad_(any(k.endswith(s) for s in TRAINABLE_SUFFIXES))
         for k, v in params.items()}
    hw = _t(np.asarray(head_w, dtype=np.float64).reshape(-1), dtype).clone().requires_grad_(True)
    hb = _t(np.asarray(head_b, dtype=np.float64).reshape(-1), dtype).clone().requires_grad_(True)
    k1, k2 = [], []
    slack = []
    rm = [None if m is None else [_t(a, dtype) for a in m] for m in relu_masks]
    ps = [None if m is None else [torch.as_tensor(np.asarray(a)).to(torch.bool) for a in m] for m in pool_selects]
    gs = [None if m is None else torch.as_tensor(np.asarray(m)) for m in gmax_selects]
    e1, s1 = _encoder_forward_train_torch(_t(x1, dtype), P, dropout_masks=dropout_masks[0], keep=k1, relu_masks=rm[0],
                                          pool_selects=ps[0], gmax_select=gs[0], slack=slack)
    e2, s2 = _encoder_forward_train_torch(_t(x2, dtype), P, dropout_masks=dropout_masks[1], keep=k2, relu_masks=rm[1],
                                          pool_selects=ps[1], gmax_select=gs[1], slack=slack)
    diff = e1 - e2
    if distance_metric == "uniform_euclidean":
        d = torch.sqrt(torch.clamp((diff * diff).sum(dim=-1, keepdim=True), min=0.0))
        z = d * hw[0] + hb[0]
    elif distance_metric == "weighted_l1":
        z = (diff.abs() * hw).sum(dim=-1, keepdim=True) + hb[0]
    else:
        raise NotImplementedError(distance_metric)
    p = torch.sigmoid(z)
    yt = _t(np.asarray(y, dtype=np.float64).reshape(-1, 1), dtype)
    if loss == "binary_crossentropy":
        pc = torch.clamp(p, 1e-7, 1 - 1e-7)
        lv = (-(yt * torch.log(pc)) - (1 - yt) * torch.log(1 - pc)).mean()
    elif loss == "contrastive_loss":
        lv = ((1 - yt) * p * p + yt * torch.clamp(1.0 - p, min=0.0) ** 2).mean()
    else:
        raise NotImplementedError(loss)
    lv.backward()
    grads = {k: v.grad.numpy().copy() for k, v in P.items() if v.requires_grad}
    return dict(loss=float(lv.item()), prob=p.detach().numpy(), grads=grads,
                head_w_grad=hw.grad.numpy().copy(), head_b_grad=hb.grad.numpy().copy(),
                stats=[[(m.numpy(), v.numpy(), n) for (m, v, n) in s] for s in (s1, s2)],
                e1=e1.detach().numpy(), e2=e2.detach().numpy(), select_slack=max(slack) if slack else 0.0,
                u=[[h.detach().numpy() for h in k] for k in (k1, k2)],
                du=[[h.grad.numpy().copy() for h in k] for k in (k1, k2)])


def classifier_train_step_grads(params, head_kernel, head_bias, x, y_onehot, dtype=torch.float64,
                                dropout_masks=None, relu_masks=None, pool_selects=None, gmax_select=None):
    """Encoder + Dense(num_classes, softmax) + categorical_crossentropy
    (experiments/train_classifier.py:110-115), training mode."""
    P = {k: _t(v, dtype).clone().requires_grad_(any(k.endswith(s) for s in TRAINABLE_SUFFIXES))
         for k, v in params.items()}
    hk = _t(head_kernel, dtype).clone().requires_grad_(True)
    hb = _t(head_bias, dtype).clone().requires_grad_(True)
    rm = None if relu_masks is None else [_t(a, dtype) for a in relu_masks]
    ps = None if pool_selects is None else [torch.as_tensor(np.asarray(a)).to(torch.bool) for a in pool_selects]
    gs = None if gmax_select is None else torch.as_tensor(np.asarray(gmax_select))
    slack = []
    emb, stats = _encoder_forward_train_torch(_t(x, dtype), P, dropout_masks=dropout_masks, relu_masks=rm,
                                              pool_selects=ps, gmax_select=gs, slack=slack)
    logits = emb @ hk + hb
    logp = torch.log_softmax(logits, dim=-1)
    lv = -(_t(y_onehot, dtype) * logp).sum(dim=-1).mean()
    lv.backward()
    grads = {k: v.grad.numpy().copy() for k, v in P.items() if v.requires_grad}
    return dict(loss=float(lv.item()), grads=grads, head_kernel_grad=hk.grad.numpy().copy(),
                head_bias_grad=hb.grad.numpy().copy(), stats=[(m.numpy(), v.numpy(), n) for (m, v, n) in stats],
                select_slack=max(slack) if slack else 0.0,
                probs=torch.softmax(logits, dim=-1).detach().numpy())
