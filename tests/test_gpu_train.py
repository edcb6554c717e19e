"""Training-path parity (GPU box): train-mode forward, every gradient, BN moving statistics and the Keras-Adam update
of one siamese / classifier step against the autograd oracle (oracle/voicemap_oracle.py, fp64)."""
import numpy as np
import pytest
import torch

from oracle import voicemap_oracle as O

pytestmark = pytest.mark.gpu


def _rel(a, ref):
    a, ref = np.asarray(a, np.float64), np.asarray(ref, np.float64)
    return np.abs(a - ref).max() / (np.abs(ref).max() + 1e-30)


def _grad_errors(grads, ref_grads):
    """max|d| per tensor relative to max(|ref tensor|, 1e-3 * the largest gradient entry of the whole step): a tensor
    whose true gradient vanishes (the embedding bias of a siamese net cancels in e1 - e2) is judged on the global
    scale instead of dividing by ~0."""
    floor = 1e-3 * max(np.abs(np.asarray(g)).max() for g in ref_grads.values())
    return {k: np.abs(np.asarray(grads[k], np.float64).reshape(np.asarray(g).shape) - g).max() /
            max(np.abs(g).max(), floor) for k, g in ref_grads.items()}


def _patterns(tr, n, branches=2):
    """The device's own discrete decisions of the last step, per branch, for the oracle: ReLU pattern, the winner of
    every MaxPool window and the winning window of GlobalMaxPool1D.  Where two candidates agree to within rounding, fp32
    and fp64 legitimately decide differently and the gradient -- which follows the winner -- differs by a whole entry
    without either side being wrong; the oracle therefore differentiates along the device's decisions and reports
    (``select_slack``) how far any of them is from a true maximum."""
    rows = [slice(br * n, (br + 1) * n) for br in range(branches)]
    relu = [[tr.relu_pattern(b)[r].cpu().numpy().astype(np.float64) for b in range(4)] for r in rows]
    pool = [[tr.argmax_flags(b)[r].cpu().numpy() for b in range(4)] for r in rows]
    gmax = [tr.jstar[r].cpu().numpy() for r in rows]
    return relu, pool, gmax


def _make(filters, emb, loss, metric="uniform_euclidean", seed=0, dropout=0.0, precision=3, bwd_precision=3):
    from voicemap_b200.keras_compat import Adam
    from voicemap_b200.models import build_siamese_net, get_baseline_convolutional_encoder
    from voicemap_b200.training import TrainEngine
    params = O.init_encoder_params(filters, emb, seed=seed, randomize_bn=False, random_bias=True)
    rng = np.random.default_rng(seed + 1)
    for i in range(1, 5):   # non-trivial gamma / beta (incl. negative gamma)
        params[f"bn{i}_gamma"] = rng.uniform(-1.2, 1.5, params[f"bn{i}_gamma"].shape).astype(np.float32)
        params[f"bn{i}_beta"] = rng.normal(0, 0.2, params[f"bn{i}_beta"].shape).astype(np.float32)
    enc = get_baseline_convolutional_encoder(filters, emb, dropout=dropout)
    enc.set_named_weights(params)
    sia = build_siamese_net(enc, (1024, 1), metric)
    if metric == "uniform_euclidean":
        sia.head_weights["head_kernel"][:] = 0.05
        sia.head_weights["head_bias"][:] = -0.3
    opt = Adam(clipnorm=1.0)
    sia.compile(loss=loss, optimizer=opt)
    tr = TrainEngine(sia, opt, sia.loss, precision=precision, bwd_precision=bwd_precision)
    return params, sia, tr


# gradient tolerance per backward arithmetic (max |error| of a tensor relative to its largest entry, fp64 autograd
# oracle; measured <= 2.2e-4 / 6.5e-4 / 8e-4 at this size, tools/train_parity_probe.py): 3 = two-plane fp16 gradients, three
# MMAs per step; 2 = one-plane gradients (every dU element rounded to 11 significant bits, unbiased) against two-plane
# activations / weights; 1 = one plane each.  The forward pass is precision 3 throughout: the siamese head
# differentiates through e1 - e2, which amplifies the fp16 + fp8 forward's 1e-5 embedding error to 1e-2 in the gradients
# of an untrained net (the loss itself stays within 1e-5).
GRAD_TOL = {3: 5e-4, 2: 2e-3, 1: 2e-3}


@pytest.mark.parametrize("precision,bwd_precision", [(3, 3), (3, 2), (3, 1)])
@pytest.mark.parametrize("filters,emb,loss,metric", [(32, 16, "binary_crossentropy", "uniform_euclidean"),
                                                     (128, 64, "contrastive_loss", "uniform_euclidean"),
                                                     (64, 32, "binary_crossentropy", "weighted_l1")])
def test_siamese_step_forward_and_gradients(filters, emb, loss, metric, precision, bwd_precision):
    params, sia, tr = _make(filters, emb, loss, metric, precision=precision, bwd_precision=bwd_precision)
    n, length = 4, 1024
    x1 = O.synthetic_clips(n, length, seed=11)
    x2 = O.synthetic_clips(n, length, seed=12)
    y = np.array([0, 0, 1, 1], dtype=np.float32)
    hw, hb = sia.head_weights["head_kernel"].reshape(-1).copy(), sia.head_weights["head_bias"].copy()
    lv, acc = tr.siamese_step(x1, x2, y, apply=False)
    torch.cuda.synchronize()
    # the oracle takes the device's discrete decisions (see _patterns): branch 1 = rows [0, n), branch 2 = [n, 2n)
    masks, pool, gmax = _patterns(tr, n)
    ref = O.siamese_train_step_grads(params, hw, hb, x1, x2, y, loss=loss, distance_metric=metric, relu_masks=masks,
                                     pool_selects=pool, gmax_selects=gmax)
    assert ref["select_slack"] < 1e-5
    # train-mode forward
    emb_gpu = tr.embv.cpu().numpy()
    assert _rel(emb_gpu[:n], ref["e1"]) < 1e-4 and _rel(emb_gpu[n:], ref["e2"]) < 1e-4
    assert abs(lv.item() - ref["loss"]) <= 1e-4 * abs(ref["loss"])
    # batch statistics per branch
    for branch in range(2):
        for b in range(4):
            m, v, _ = ref["stats"][branch][b]
            bnc = tr.bnc[b][branch].cpu().numpy()
            assert _rel(bnc[:, 2], m) < 1e-4
            assert _rel(1.0 / np.square(bnc[:, 3]) - O.BN_EPS, v) < 1e-3
    # gradients
    # kernel-level diagnostics: block-1 activations and their gradients (the last dU left in the buffer is block 1's)
    u1 = tr.activation(0).cpu().numpy()            # fp16 copy kept for the backward pass: 2^-11 relative
    u1_ref = np.concatenate([ref["u"][0][0], ref["u"][1][0]], axis=0)
    du1 = tr.block_gradient(0).cpu().numpy()
    # the oracle keeps d loss / d u of the post-ReLU tensor; the kernel stores the gradient of the conv output
    du1_ref = np.concatenate([ref["du"][0][0], ref["du"][1][0]], axis=0) * (u1 > 0)
    err = np.abs(du1 - du1_ref)
    bad = np.argwhere(err > 1e-4 * np.abs(du1_ref).max())
    print("U1 rel err", _rel(u1, u1_ref), "dU1 rel err", err.max() / np.abs(du1_ref).max(),
          "bad entries", len(bad), "of", err.size)
    if len(bad):
        print("  bad clips", np.unique(bad[:, 0]), "bad l (first 12)", np.unique(bad[:, 1])[:12], "n bad l",
              len(np.unique(bad[:, 1])), "bad c", np.unique(bad[:, 2])[:16], "n bad c", len(np.unique(bad[:, 2])))
        i = tuple(bad[0])
        print("  first bad", i, "got", du1[i], "ref", du1_ref[i], "u", u1[i])
    grads = tr.gradients()
    refg = dict(ref["grads"], head_kernel=ref["head_w_grad"], head_bias=ref["head_b_grad"])
    worst = _grad_errors(grads, refg)
    print({k: f"{v:.2e}" for k, v in worst.items()})
    assert max(worst.values()) < GRAD_TOL[bwd_precision], worst


def test_moving_statistics_and_adam_update():
    params, sia, tr = _make(32, 16, "binary_crossentropy")
    n, length = 4, 1024
    x1, x2 = O.synthetic_clips(n, length, seed=21), O.synthetic_clips(n, length, seed=22)
    y = np.array([0, 1, 0, 1], dtype=np.float32)
    hw, hb = sia.head_weights["head_kernel"].reshape(-1).copy(), sia.head_weights["head_bias"].copy()
    ref = O.siamese_train_step_grads(params, hw, hb, x1, x2, y)
    before = {k: v.clone() for k, v in tr.p.items()}
    tr.siamese_step(x1, x2, y, apply=True)
    torch.cuda.synchronize()
    # moving statistics: two sequential Keras updates (branch 1 then branch 2)
    for b in range(4):
        mm, mv = params[f"bn{b + 1}_mean"].astype(np.float64), params[f"bn{b + 1}_var"].astype(np.float64)
        for branch in range(2):
            m, v, cnt = ref["stats"][branch][b]
            mm, mv = O.bn_moving_update(mm, mv, m, v, cnt)
        assert _rel(tr.moving[f"bn{b + 1}_mean"].cpu().numpy(), mm) < 1e-4
        assert _rel(tr.moving[f"bn{b + 1}_var"].cpu().numpy(), mv) < 1e-4
    # Adam(clipnorm=1) arithmetic: oracle update applied to the SAME gradients the device produced (gradient parity
    # is checked above; the first Adam step is sign-like, so it must not be compared across gradient sources)
    p = {k: before[k].cpu().numpy().astype(np.float64) for k in before}
    g = {k: v.astype(np.float64) for k, v in tr.gradients().items()}
    m0 = {k: np.zeros_like(v) for k, v in p.items()}
    v0 = {k: np.zeros_like(v) for k, v in p.items()}
    O.keras_adam_step(p, g, m0, v0, t=1, clipnorm=1.0)
    for k in p:
        step_ref = p[k] - before[k].cpu().numpy()
        step_gpu = tr.p[k].cpu().numpy() - before[k].cpu().numpy()
        assert np.abs(step_gpu - step_ref).max() <= 1e-3 * np.abs(step_ref).max() + 1e-8, k  # lr = 1e-3
    # second step exercises the moment recursion and the bias-corrected learning rate
    tr.siamese_step(x1, x2, y, apply=False)
    g2 = {k: v.astype(np.float64) for k, v in tr.gradients().items()}
    prev = {k: v.clone() for k, v in tr.p.items()}
    tr.apply_gradients()
    torch.cuda.synchronize()
    p2 = {k: prev[k].cpu().numpy().astype(np.float64) for k in prev}
    O.keras_adam_step(p2, g2, m0, v0, t=2, clipnorm=1.0)
    for k in p2:
        step_ref = p2[k] - prev[k].cpu().numpy()
        step_gpu = tr.p[k].cpu().numpy() - prev[k].cpu().numpy()
        assert np.abs(step_gpu - step_ref).max() <= 1e-3 * np.abs(step_ref).max() + 1e-8, k  # lr = 1e-3


def test_classifier_step_gradients():
    from voicemap_b200.keras_compat import Adam, Dense
    from voicemap_b200.models import get_baseline_convolutional_encoder
    from voicemap_b200.training import TrainEngine
    filters, emb, classes, n, length = 32, 16, 7, 6, 1024
    params = O.init_encoder_params(filters, emb, seed=5, random_bias=True)
    clf = get_baseline_convolutional_encoder(filters, emb, (length, 1), dropout=0.0)
    clf.set_named_weights(params)
    clf.add(Dense(classes, activation="softmax"))
    opt = Adam(clipnorm=1.0)
    clf.compile(loss="categorical_crossentropy", optimizer=opt, metrics=["accuracy"])
    tr = TrainEngine(clf, opt, clf.loss)
    x = O.synthetic_clips(n, length, seed=31)
    y = np.eye(classes, dtype=np.float32)[np.arange(n) % classes]
    lv, acc = tr.classifier_step(x, y, apply=False)
    torch.cuda.synchronize()
    masks, pool, gmax = _patterns(tr, n, branches=1)
    ref = O.classifier_train_step_grads(params, clf.weights["head_kernel"], clf.weights["head_bias"], x, y,
                                        relu_masks=masks[0], pool_selects=pool[0], gmax_select=gmax[0])
    assert ref["select_slack"] < 1e-5
    assert abs(lv.item() - ref["loss"]) <= 1e-4 * abs(ref["loss"])
    grads = tr.gradients()
    worst = _grad_errors(grads, dict(ref["grads"], head_kernel=ref["head_kernel_grad"], head_bias=ref["head_bias_grad"]))
    print({k: f"{v:.2e}" for k, v in worst.items()})
    assert max(worst.values()) < 2e-3, worst


def test_dropout_mask_semantics():
    """SpatialDropout1D: one keep/drop decision per (clip, channel), scaled by 1/(1-p); explicit masks make the
    train-mode forward comparable with the oracle."""
    params, sia, tr = _make(32, 16, "binary_crossentropy", dropout=0.25)
    n, length = 4, 1024
    x1, x2 = O.synthetic_clips(n, length, seed=41), O.synthetic_clips(n, length, seed=42)
    y = np.array([0, 0, 1, 1], dtype=np.float32)
    rng = np.random.default_rng(0)
    masks = [(rng.random((2 * n, c)) < 0.75).astype(np.float32) / 0.75 for c in tr.channels]
    dm = [torch.from_numpy(m).cuda() for m in masks]
    om1 = [torch.from_numpy(m[:n, None, :]).double() for m in masks]
    om2 = [torch.from_numpy(m[n:, None, :]).double() for m in masks]
    hw, hb = sia.head_weights["head_kernel"].reshape(-1).copy(), sia.head_weights["head_bias"].copy()
    lv, _ = tr.siamese_step(x1, x2, y, apply=False, masks=dm)
    torch.cuda.synchronize()
    rmasks, pool, gmax = _patterns(tr, n)
    ref = O.siamese_train_step_grads(params, hw, hb, x1, x2, y, dropout_masks=(om1, om2), relu_masks=rmasks,
                                     pool_selects=pool, gmax_selects=gmax)
    assert ref["select_slack"] < 1e-5
    assert abs(lv.item() - ref["loss"]) <= 1e-4 * abs(ref["loss"])
    assert max(_grad_errors(tr.gradients(), ref["grads"]).values()) < 2e-3
    assert tr._set_masks(8, None)
    m = tr.masks[0].cpu().numpy()
    assert set(np.unique(m)).issubset({0.0, np.float32(1 / 0.75)}) and m.shape == (8, 32)


def test_fit_generator_reduces_loss(tmp_path):
    """A few epochs on a separable toy problem through the Keras-style API incl. callbacks."""
    from voicemap_b200.keras_compat import Adam, CSVLogger
    from voicemap_b200.models import build_siamese_net, get_baseline_convolutional_encoder
    rng = np.random.default_rng(0)
    length, n = 1024, 16

    def gen():
        while True:
            t = np.arange(length)[None, :]
            f1 = rng.uniform(0.02, 0.04, (n, 1)); f2 = rng.uniform(0.2, 0.3, (n, 1))
            cls_a = rng.integers(0, 2, n); same = (np.arange(n) < n // 2)
            cls_b = np.where(same, cls_a, 1 - cls_a)
            fa = np.where(cls_a[:, None] == 0, f1, f2); fb = np.where(cls_b[:, None] == 0, f1, f2)
            xa = 0.05 * np.sin(2 * np.pi * fa * t + rng.uniform(0, 6, (n, 1)))
            xb = 0.05 * np.sin(2 * np.pi * fb * t + rng.uniform(0, 6, (n, 1)))
            yield [xa[:, :, None], xb[:, :, None]], (~same).astype(np.float64)[:, None]

    enc = get_baseline_convolutional_encoder(16, 8, dropout=0.0)
    sia = build_siamese_net(enc, (length, 1))
    sia.compile(loss="binary_crossentropy", optimizer=Adam(lr=3e-3, clipnorm=1.0), metrics=["accuracy"])
    hist = sia.fit_generator(gen(), steps_per_epoch=15, epochs=4, validation_data=gen(), validation_steps=2, verbose=0,
                             callbacks=[CSVLogger(str(tmp_path / "log.csv"))])
    assert hist[-1]["loss"] < hist[0]["loss"]
    assert "val_loss" in hist[-1] and "acc" in hist[-1]
    assert np.isfinite(sia.predict(next(gen())[0])).all()


def test_sync_bn_split_path_equals_fused_path_on_one_rank():
    """Synchronised BatchNorm splits the statistics at the all-reduce point (vm_bn_stats_sums / _from_sums,
    vm_bn_bwd_sums / _from_sums).  With one rank (identity all-reduce) the split path must reproduce the fused
    kernels bit for bit: loss, every gradient, moving statistics."""
    n, length = 4, 1024
    x1, x2 = O.synthetic_clips(n, length, seed=21), O.synthetic_clips(n, length, seed=22)
    y = np.array([0, 1, 0, 1], dtype=np.float32)
    out = []
    for sync in (False, True):
        _, sia, tr = _make(64, 32, "contrastive_loss", seed=3)
        calls = []
        if sync:
            tr.set_sync_bn(lambda t: calls.append(t.numel()), 1)
        lv, _ = tr.siamese_step(x1, x2, y, apply=False)
        torch.cuda.synchronize()
        out.append((float(lv.item()), tr.gradients(), {k: v.clone() for k, v in tr.moving.items()}))
        if sync:   # 4 blocks forward + 4 backward, each (2 groups, C, 2) doubles
            assert calls == [2 * c * 2 for c in (64, 128, 192, 256)] + [2 * c * 2 for c in (256, 192, 128, 64)]
    assert out[0][0] == out[1][0]
    for k in out[0][1]:
        assert np.array_equal(out[0][1][k], out[1][1][k]), k
    for k in out[0][2]:
        assert torch.equal(out[0][2][k], out[1][2][k]), k


def test_train_step_at_baseline_shape_64_pairs_12000():
    """BASELINE config[2]'s per-step shape on one GPU: 64 pairs (128 clips) x 12000 samples, filters 128, embedding 64,
    contrastive loss (experiments/siamese_contrastive_loss.py:70,76-83) against the fp64 autograd oracle: train-mode
    embeddings and loss at 1e-4, batch statistics, and every gradient tensor at the tolerance of its backward
    arithmetic (GRAD_TOL).  Both backward modes share one oracle evaluation: they run the same forward pass, so the
    discrete decisions handed to the oracle (ReLU pattern, pool winners) are the same (checked)."""
    n, length, filters, emb = 64, 12000, 128, 64
    x1 = O.synthetic_clips(n, length, seed=111)
    x2 = O.synthetic_clips(n, length, seed=112)
    y = (np.arange(n) >= n // 2).astype(np.float32)
    ref, pattern = None, None
    for bwd in (1, 3):
        params, sia, tr = _make(filters, emb, "contrastive_loss", seed=7, bwd_precision=bwd)
        hw, hb = sia.head_weights["head_kernel"].reshape(-1).copy(), sia.head_weights["head_bias"].copy()
        lv, _ = tr.siamese_step(x1, x2, y, apply=False)
        torch.cuda.synchronize()
        masks, pool, gmax = _patterns(tr, n)
        if ref is None:
            pattern = (masks, pool, gmax)
            ref = O.siamese_train_step_grads(params, hw, hb, x1, x2, y, loss="contrastive_loss", relu_masks=masks,
                                             pool_selects=pool, gmax_selects=gmax)
            assert ref["select_slack"] < 1e-5      # every device winner is a maximum up to rounding
        else:
            for got, first in zip((masks, pool), pattern[:2]):
                assert all(np.array_equal(a, b) for sa, sb in zip(got, first) for a, b in zip(sa, sb))
            assert all(np.array_equal(a, b) for a, b in zip(gmax, pattern[2]))
        emb_gpu = tr.embv.cpu().numpy()
        assert _rel(emb_gpu[:n], ref["e1"]) < 1e-4 and _rel(emb_gpu[n:], ref["e2"]) < 1e-4
        assert abs(lv.item() - ref["loss"]) <= 1e-4 * abs(ref["loss"])
        for branch in range(2):
            for b in range(4):
                m, v, _ = ref["stats"][branch][b]
                bnc = tr.bnc[b][branch].cpu().numpy()
                assert _rel(bnc[:, 2], m) < 1e-4
                assert _rel(1.0 / np.square(bnc[:, 3]) - O.BN_EPS, v) < 1e-3
        refg = dict(ref["grads"], head_kernel=ref["head_w_grad"], head_bias=ref["head_b_grad"])
        worst = _grad_errors(tr.gradients(), refg)
        print(f"64 pairs x 12000, bwd_precision {bwd}: loss rel err {abs(lv.item() - ref['loss']) / abs(ref['loss']):.2e}",
              {k: f"{v:.2e}" for k, v in worst.items()})
        assert max(worst.values()) < GRAD_TOL[bwd], (bwd, worst)
        del tr, sia
        torch.cuda.empty_cache()


@pytest.mark.parametrize("n,c,e", [(5, 512, 64), (64, 512, 64), (7, 136, 24), (3, 64, 100)])
@pytest.mark.parametrize("metric,loss_id", [(0, 1), (0, 2), (1, 2)])
def test_siamese_head_train_equals_the_four_separate_calls(n, c, e, metric, loss_id):
    """vm_siamese_head_train (the siamese training head in two launches) against vm_dense_fwd + vm_pair_head_loss_fwd +
    vm_pair_head_loss_bwd + vm_dense_bwd, which it replaces in the training step: embeddings, probabilities, d_emb,
    d_gmax bit for bit (same summation orders, or exact negation); loss, accuracy and the reduced gradients to float
    rounding of differently ordered sums."""
    import ctypes as C
    from voicemap_b200 import _lib
    lib = _lib.load()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    g = torch.Generator().manual_seed(100 * n + c + e)
    dev = "cuda"
    gmax = torch.randn(2 * n, c, generator=g).to(dev)
    w = (torch.randn(c, e, generator=g) / np.sqrt(c)).to(dev)
    b = (0.1 * torch.randn(e, generator=g)).to(dev)
    hw = (torch.rand(1 if metric == 0 else e, generator=g) * 0.5 + 0.1).to(dev)
    hb = torch.tensor([-0.4]).to(dev)
    y = (torch.arange(n) % 2).to(torch.float32).to(dev)
    scale = 8.0
    p = lambda t: C.c_void_p(t.data_ptr())   # noqa: E731
    # separate calls
    emb0 = torch.empty(2 * n, e, device=dev)
    prob0, loss0, acc0 = torch.empty(n, device=dev), torch.zeros(1, device=dev), torch.zeros(1, device=dev)
    d_emb0, d_gmax0 = torch.empty(2 * n, e, device=dev), torch.empty(2 * n, c, device=dev)
    dw0, db0 = torch.empty(c, e, device=dev), torch.empty(e, device=dev)
    dhw0, dhb0 = torch.empty_like(hw), torch.empty(1, device=dev)
    _lib.check(lib.vm_dense_fwd(p(gmax), 2 * n, c, p(w), p(b), e, p(emb0), st), "dense_fwd")
    _lib.check(lib.vm_pair_head_loss_fwd(p(emb0[:n]), p(emb0[n:]), n, e, metric, p(hw), p(hb), p(y), loss_id, None,
                                         p(prob0), p(loss0), st), "head_fwd")
    _lib.check(lib.vm_pair_head_loss_bwd(p(emb0), n, e, metric, p(hw), p(hb), p(y), loss_id, C.c_float(scale), p(d_emb0),
                                         p(dhw0), p(dhb0), p(acc0), st), "head_bwd")
    _lib.check(lib.vm_dense_bwd(p(gmax), p(d_emb0), p(w), 2 * n, c, e, p(dw0), p(db0), p(d_gmax0), st), "dense_bwd")
    # fused call
    emb1 = torch.empty(2 * n, e, device=dev)
    prob1, la1 = torch.empty(n, device=dev), torch.zeros(2, device=dev)
    d_emb1, d_gmax1 = torch.empty(2 * n, e, device=dev), torch.empty(2 * n, c, device=dev)
    dw1, db1 = torch.empty(c, e, device=dev), torch.empty(e, device=dev)
    dhw1, dhb1 = torch.empty_like(hw), torch.empty(1, device=dev)
    rec = torch.empty(n, 4, device=dev)
    _lib.check(lib.vm_siamese_head_train(p(gmax), n, c, e, p(w), p(b), metric, p(hw), p(hb), p(y), loss_id,
                                         C.c_float(scale), p(emb1), p(prob1), p(d_emb1), p(d_gmax1), p(rec), p(dw1),
                                         p(db1), p(dhw1), p(dhb1), p(la1), st), "siamese_head_train")
    torch.cuda.synchronize()
    assert torch.equal(emb1, emb0) and torch.equal(prob1, prob0) and torch.equal(d_emb1, d_emb0)

    def close(a, ref, tol=2e-6):
        a, ref = a.double().cpu().numpy(), ref.double().cpu().numpy()
        return np.abs(a - ref).max() <= tol * max(np.abs(ref).max(), 1e-30) + 1e-12
    assert close(d_gmax1, d_gmax0) and close(dw1, dw0) and close(dhw1, dhw0) and close(dhb1, dhb0)
    assert np.abs(db1.cpu().numpy() - db0.cpu().numpy()).max() <= 1e-6 * np.abs(d_emb0.cpu().numpy()).max() * n
    assert close(la1[:1], loss0) and float(la1[1]) == float(acc0[0])
