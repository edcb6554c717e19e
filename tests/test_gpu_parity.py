"""Parity tests proper (GPU box): the CUDA path, called through the C ABI (ctypes), against the CPU oracle on the
same seeded inputs, against the committed golden fixture, and -- at BASELINE.json's full batch -- through
size-independent properties.  Tolerance (north_star / SURVEY.md 8(d)): per-clip ||e_gpu - e_ref|| / ||e_ref|| <= 1e-4
and max|d| / max|e_ref| <= 1e-4 against the fp32 oracle in both parity modes: precision 3 (fp16 x 3, measured ~7e-6)
and precision 2 (fp16 + one e5m2-pair fp8 correction product, the default of predict(); measured 1e-5 .. 4.5e-5).
Per-block tests run precision 3 at fp32-grade tolerances, plus precision 2 at its own (the merged view of a
precision-2 plane pair carries ~14 bits)."""
import os

import numpy as np
import pytest
import torch

from oracle import voicemap_oracle as O

pytestmark = pytest.mark.gpu

TOL = 1e-4
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "encoder_f128_e64.npz")


def _engine(filters, emb, params, precision=3):
    from voicemap_b200.engine import EncoderEngine
    eng = EncoderEngine(filters, emb, precision=precision)
    eng.set_weights(params)
    return eng


def _rel(a, ref):
    a, ref = np.asarray(a, np.float64), np.asarray(ref, np.float64)
    return np.abs(a - ref).max() / np.abs(ref).max()


def _per_clip(a, ref):
    a, ref = np.asarray(a, np.float64), np.asarray(ref, np.float64)
    return (np.linalg.norm(a - ref, axis=1) / np.linalg.norm(ref, axis=1)).max()


def _block_ref(x_in, params, b, dtype=torch.float32):
    h = O._t(x_in, dtype)
    h = O.conv1d_same_relu(h, O._t(params[f"conv{b}_kernel"], dtype), O._t(params[f"conv{b}_bias"], dtype))
    h = O.batchnorm_eval(h, O._t(params[f"bn{b}_gamma"], dtype), O._t(params[f"bn{b}_beta"], dtype),
                         O._t(params[f"bn{b}_mean"], dtype), O._t(params[f"bn{b}_var"], dtype))
    return O.maxpool1d_valid(h, O.POOLS[b - 1]).numpy()


@pytest.fixture(scope="module")
def stress_params():
    # Keras-init kernels, random biases, randomised BN incl. negative gamma and tiny variances
    return O.init_encoder_params(128, 64, seed=0, randomize_bn=True, random_bias=True)


@pytest.mark.parametrize("n,length", [(2, 2048), (3, 1999), (1, 256), (2, 261), (5, 32), (1, 12000)])
def test_block1_parity_incl_ragged_lengths(stress_params, n, length):
    eng = _engine(128, 64, stress_params)
    x = O.synthetic_clips(n, length, seed=5)
    hi, lo = eng.block1(torch.from_numpy(x[:, :, 0].copy()).cuda())
    got = eng.merge_planes(hi, lo).cpu().numpy()
    ref = _block_ref(x, stress_params, 1)
    assert got.shape == ref.shape == (n, length // 4, 128)
    assert _rel(got, ref) < 1e-5


@pytest.mark.parametrize("block,n,length", [(2, 2, 512), (2, 3, 499), (2, 1, 2), (2, 2, 257), (3, 2, 300), (3, 1, 1500),
                                            (4, 2, 750), (4, 3, 187)])
def test_block234_parity_incl_ragged_lengths(stress_params, block, n, length):
    eng = _engine(128, 64, stress_params)
    cin = 128 * (block - 1)
    rng = np.random.default_rng(block * 100 + length)
    x = (rng.normal(0, 1.0, (n, length, cin)) * rng.uniform(0.1, 3.0, (1, 1, cin))).astype(np.float32)
    hi, lo = eng.split_planes(torch.from_numpy(x).cuda())
    ref = _block_ref(x, stress_params, block, torch.float64)
    if block < 4:
        oh, ol = eng.block3(block, hi, lo)
        got = eng.merge_planes(oh, ol).cpu().numpy()
        assert got.shape == ref.shape == (n, length // 2, 128 * block)
        assert _rel(got, ref) < 2e-5
    else:
        part = eng.block3(4, hi, lo, gmax=True)
        emb, g = eng.gmax_dense(part, with_gmax=True)
        assert _rel(g.cpu().numpy(), ref.max(axis=1)) < 2e-5


@pytest.mark.parametrize("block,n,length", [(2, 2, 512), (2, 3, 499), (3, 2, 300), (4, 2, 750), (4, 3, 187)])
def test_block234_precision2_incl_ragged_lengths(stress_params, block, n, length):
    """precision 2: Xh*Wh on the fp16 pipe + (Xl*Wh + Xh*Wl) as one fp8 product over e5m2 byte pairs."""
    eng = _engine(128, 64, stress_params, precision=2)
    cin = 128 * (block - 1)
    rng = np.random.default_rng(block * 100 + length)
    x = (rng.normal(0, 1.0, (n, length, cin)) * rng.uniform(0.1, 3.0, (1, 1, cin))).astype(np.float32)
    hi, q = eng.split_planes(torch.from_numpy(x).cuda())
    ref = _block_ref(x, stress_params, block, torch.float64)
    if block < 4:
        oh, oq = eng.block3(block, hi, q)
        got = eng.merge_planes(oh, oq).cpu().numpy()
        assert got.shape == ref.shape == (n, length // 2, 128 * block)
        assert _rel(got, ref) < 1e-4
    else:
        part = eng.block3(4, hi, q, gmax=True)
        emb, g = eng.gmax_dense(part, with_gmax=True)
        assert _rel(g.cpu().numpy(), ref.max(axis=1)) < 1e-4


def test_precision2_plane_pair_semantics():
    """(hi, Q): hi = fp16(x); Q = {e5m2(x / 64), e5m2((x - hi) * 64)} (upper, lower byte)."""
    from voicemap_b200.engine import EncoderEngine
    eng = EncoderEngine(128, 64, precision=2)
    x = torch.randn(1 << 14, device="cuda") * torch.logspace(-2, 3, 1 << 14, device="cuda")
    hi, q = eng.split_planes(x)
    assert torch.equal(hi, x.to(torch.float16))
    qb = q.view(torch.uint8).reshape(-1, 2)                      # little endian: [lower, upper]
    lower = qb[:, 0].contiguous().view(torch.float8_e5m2).float()
    upper = qb[:, 1].contiguous().view(torch.float8_e5m2).float()
    resid = (x - hi.float()) * 64.0
    assert ((lower - resid).abs() <= 0.126 * resid.abs() + 2.0 ** -17).all()
    assert ((upper - x / 64.0).abs() <= 0.126 * (x / 64.0).abs() + 2.0 ** -17).all()
    back = eng.merge_planes(hi, q)
    assert ((back - x).abs() <= 2.0 ** -13 * x.abs() + 2.0 ** -22).all()


def test_planes_roundtrip_is_fp32_grade():
    from voicemap_b200.engine import EncoderEngine
    eng = EncoderEngine(128, 64)
    x = torch.randn(1 << 16, device="cuda") * torch.logspace(-3, 3, 1 << 16, device="cuda")
    hi, lo = eng.split_planes(x)
    back = eng.merge_planes(hi, lo)
    rel = ((back - x).abs() / x.abs().clamp_min(2.0 ** -3)).max().item()
    assert rel < 2.0 ** -21


@pytest.mark.parametrize("precision", [2, 3])
@pytest.mark.parametrize("n,length,padded", [(8, 12000, False), (8, 12000, True), (5, 11999, False), (3, 6000, True),
                                             (2, 48000, False), (4, 4000, False)])
def test_encoder_parity_config0_and_lengths(stress_params, n, length, padded, precision):
    """BASELINE config[0] (batch 8, 3 s) plus the reference's n_seconds sweep lengths and a raw 16 kHz clip."""
    eng = _engine(128, 64, stress_params, precision)
    x = O.synthetic_clips(n, length, seed=1234, padded=padded)
    got = eng.forward(torch.from_numpy(x[:, :, 0].copy()).cuda()).cpu().numpy()
    ref32 = O.encoder_forward(x, stress_params, torch.float32)
    ref64 = O.encoder_forward(x, stress_params, torch.float64)
    assert _per_clip(got, ref32) <= TOL and _rel(got, ref32) <= TOL
    assert _per_clip(got, ref64) <= TOL


@pytest.mark.parametrize("precision", [2, 3])
def test_encoder_parity_keras_default_init(precision):
    params = O.init_encoder_params(128, 64, seed=7)  # gamma 1, beta 0, mean 0, var 1, zero biases
    eng = _engine(128, 64, params, precision)
    x = O.synthetic_clips(4, 12000, seed=99)
    got = eng.forward(torch.from_numpy(x[:, :, 0].copy()).cuda()).cpu().numpy()
    ref = O.encoder_forward(x, params, torch.float32)
    assert _per_clip(got, ref) <= TOL


def test_golden_fixture():
    z = np.load(GOLDEN)
    params = O.init_encoder_params(int(z["filters"]), int(z["emb"]), seed=int(z["param_seed"]), randomize_bn=True,
                                   random_bias=True)
    eng = _engine(int(z["filters"]), int(z["emb"]), params)
    x = torch.from_numpy(z["x"][:, :, 0].copy()).cuda()
    got = eng.forward(x).cpu().numpy()
    assert _per_clip(got, z["emb64"]) <= TOL and _per_clip(got, z["emb32"]) <= TOL
    got2 = _engine(int(z["filters"]), int(z["emb"]), params, precision=2).forward(x).cpu().numpy()
    assert _per_clip(got2, z["emb64"]) <= TOL and _per_clip(got2, z["emb32"]) <= TOL
    hi, lo = eng.block1(x)
    b1 = eng.merge_planes(hi, lo).cpu().numpy()
    assert _rel(b1[:, :40, :], z["block1_sample"]) < 1e-5
    h2, l2 = eng.block3(2, hi, lo)
    b2 = eng.merge_planes(h2, l2).cpu().numpy()
    assert _rel(b2[:, :20, :], z["block2_sample"]) < 2e-5
    # siamese head + losses on the golden embeddings
    from voicemap_b200.engine import pair_head_loss
    e = torch.from_numpy(z["emb64"].astype(np.float32)).cuda()
    w = torch.tensor([float(z["head_w"])], device="cuda")
    b = torch.tensor([float(z["head_b"])], device="cuda")
    y = torch.from_numpy(z["y"].astype(np.float32)).cuda()
    for kind, key in (("contrastive", "contrastive64"), ("binary_crossentropy", "bce64")):
        prob, dist, loss = pair_head_loss(e[:3].contiguous(), e[3:].contiguous(), w, b, "uniform_euclidean", y, kind)
        assert np.abs(prob.cpu().numpy() - z["prob64"]).max() < 1e-5
        assert _rel(dist.cpu().numpy(), z["dist64"]) < 1e-5
        assert abs(loss.item() - float(z[key])) <= 1e-4 * abs(float(z[key]))


def test_throughput_mode_measured_tolerance(stress_params):
    """precision=1 (single fp16 MMA per MAC) is NOT parity mode; its error is measured here and bounded loosely."""
    eng = _engine(128, 64, stress_params, precision=1)
    x = O.synthetic_clips(8, 12000, seed=1234)
    got = eng.forward(torch.from_numpy(x[:, :, 0].copy()).cuda()).cpu().numpy()
    ref = O.encoder_forward(x, stress_params, torch.float64)
    err = _per_clip(got, ref)
    print(f"throughput-mode per-clip rel err: {err:.3e}")
    assert 1e-6 < err < 1e-2


def test_bce_clip_follows_keras_fp32_semantics():
    """keras binary_crossentropy clips p to [1e-7, 1 - 1e-7] in float32: saturated pairs contribute
    -log(fp32(1e-7)) resp. -log(1 - fp32(1 - 1e-7)); the oracle evaluated on float32 inputs does the same."""
    from voicemap_b200.engine import pair_head_loss
    e1 = torch.zeros(4, 8, device="cuda")
    e2 = torch.zeros(4, 8, device="cuda")
    e2[:2] += 100.0                                   # far apart -> p == 1.0 in fp32
    w = torch.tensor([1.0], device="cuda")
    b = torch.tensor([-40.0], device="cuda")          # identical pairs -> p == sigmoid(-40) ~ 4e-18
    y = torch.tensor([[0.0], [1.0], [0.0], [1.0]], device="cuda")
    prob, _, loss = pair_head_loss(e1, e2, w, b, "uniform_euclidean", y, "binary_crossentropy")
    ref = O.binary_crossentropy(y.cpu().numpy().astype(np.float32), prob.cpu().numpy().astype(np.float32))
    assert abs(loss.item() - float(ref)) <= 1e-5 * abs(float(ref))


def test_weighted_l1_head_matches_oracle():
    from voicemap_b200.engine import pair_head_loss
    rng = np.random.default_rng(3)
    e1, e2 = rng.normal(size=(7, 64)).astype(np.float32), rng.normal(size=(7, 64)).astype(np.float32)
    w, b = rng.normal(size=(64,)).astype(np.float32), np.float32(0.3)
    prob, _, _ = pair_head_loss(torch.from_numpy(e1).cuda(), torch.from_numpy(e2).cuda(), torch.from_numpy(w).cuda(),
                                torch.tensor([b], device="cuda"), "weighted_l1")
    ref, _ = O.siamese_head(e1.astype(np.float64), e2.astype(np.float64), w.astype(np.float64), float(b), "weighted_l1")
    assert np.abs(prob.cpu().numpy() - ref).max() < 1e-5


def test_models_api_predict_matches_oracle(stress_params):
    from voicemap_b200.keras_compat import Dense
    from voicemap_b200.models import build_siamese_net, get_baseline_convolutional_encoder
    enc = get_baseline_convolutional_encoder(128, 64, dropout=0.0)
    enc.set_named_weights(stress_params)
    x = O.synthetic_clips(6, 4000, seed=21).astype(np.float64)  # the reference hands float64 to predict()
    ref = O.encoder_forward(x, stress_params, torch.float32)
    assert _per_clip(enc.predict(x), ref) <= TOL
    sia = build_siamese_net(enc, (4000, 1))
    prob = sia.predict([x[:3], x[3:]])
    w, b = sia.head_weights["head_kernel"].reshape(-1)[0], sia.head_weights["head_bias"][0]
    refp, _ = O.siamese_head(ref[:3].astype(np.float64), ref[3:].astype(np.float64), float(w), float(b))
    assert prob.shape == (3, 1) and np.abs(prob - refp).max() < 1e-4
    assert sia.layers[2].predict(x).shape == (6, 64)           # voicemap/utils.py:141 path
    with pytest.raises(ValueError):
        sia.predict([x[:3], x[3:, :2000]])
    clf = get_baseline_convolutional_encoder(128, 64, (4000, 1))
    clf.set_named_weights(stress_params)
    clf.add(Dense(10, activation="softmax"))
    p = clf.predict(x)
    assert p.shape == (6, 10) and np.allclose(p.sum(axis=1), 1.0, atol=1e-5)
    with pytest.raises(ValueError):
        clf.predict(x[:, :3000])


def test_predict_pipelined_host_batch_is_bit_identical(stress_params):
    """predict() of a pinned host batch runs as growing chunks whose copies overlap the kernels
    (models._pipeline_plan); eval-mode embeddings do not depend on the chunking, bit for bit."""
    from voicemap_b200.models import _pipeline_plan, get_baseline_convolutional_encoder
    enc = get_baseline_convolutional_encoder(128, 64, dropout=0.0)
    enc.set_named_weights(stress_params)
    g = torch.Generator().manual_seed(11)
    x = (O.WHITEN_RMS * torch.randn(300, 4000, 1, generator=g)).pin_memory()
    assert len(_pipeline_plan(300)) >= 2
    got = enc.predict(x)
    eng = enc._get_engine()
    whole = eng.forward(x[:, :, 0].cuda().contiguous()).cpu().numpy()
    assert np.array_equal(got, whole)
    assert np.array_equal(enc.predict(x.numpy()), whole)        # pageable numpy input: same numbers
    assert np.array_equal(enc.predict(x), whole)                # buffer reuse across calls


def test_predict_numpy_float64_is_staged_and_bit_identical(stress_params, monkeypatch):
    """predict() of a numpy float64 batch (what the batcher and the reference's preprocessing hand over) casts chunk by
    chunk on worker threads into pinned slots (models._HostStage): same bits as the float32 batch embedded in one
    launch, with slot recycling, with a staging buffer smaller than the batch, and through the siamese model."""
    from voicemap_b200 import models as M
    enc = M.get_baseline_convolutional_encoder(128, 64, dropout=0.0)
    enc.set_named_weights(stress_params)
    rng = np.random.default_rng(5)
    x = O.WHITEN_RMS * rng.standard_normal((700, 2000, 1))       # float64; 700 clips = 6 chunks of 128 > 4 slots
    eng = enc._get_engine()
    whole = eng.forward(torch.from_numpy(x[:, :, 0].astype(np.float32)).cuda()).cpu().numpy()
    assert np.array_equal(enc.predict(x), whole)
    assert np.array_equal(enc.predict(x[::-1][::-1]), whole)     # non-contiguous view, second call (slots reused)
    assert np.array_equal(enc.predict(x[:5]), whole[:5])         # n-shot sized batch: a single chunk
    monkeypatch.setattr(M, "_PIPELINE_BUFFER_BYTES", 4 * 2000 * 256)   # device staging buffer of 256 clips: 3 passes
    enc._copy_buf = None
    assert np.array_equal(enc.predict(x), whole)
    monkeypatch.undo()
    sia = M.build_siamese_net(enc, (2000, 1))
    from_numpy = sia.predict([x[:350], x[350:]])
    from_tensor = sia.predict([torch.from_numpy(x[:350].astype(np.float32)), torch.from_numpy(x[350:].astype(np.float32))])
    assert from_numpy.shape == (350, 1) and np.array_equal(from_numpy, from_tensor)


@pytest.mark.parametrize("precision", [2, 3])
def test_full_batch_properties(stress_params, precision):
    """BASELINE config[1] size (256 clips x 12000): every one of the 256 clips against the oracle at the north-star
    tolerance, batch-composition independence (bit-exact), permutation equivariance, run-to-run determinism."""
    eng = _engine(128, 64, stress_params, precision)
    g = torch.Generator().manual_seed(5)
    x = (O.WHITEN_RMS * torch.randn(256, 12000, generator=g)).cuda()
    full = eng.forward(x).clone()
    assert torch.isfinite(full).all()
    sub = eng.forward(x[100:108].contiguous()).clone()
    assert torch.equal(sub, full[100:108])                      # a clip's embedding does not depend on its batch
    perm = torch.randperm(256, generator=g).cuda()
    assert torch.equal(eng.forward(x[perm].contiguous()), full[perm])
    ref = O.encoder_forward(x.cpu().numpy()[:, :, None], stress_params, torch.float32)      # all 256 clips
    assert ref.shape == (256, 64) and _per_clip(full.cpu().numpy(), ref) <= TOL
    # 'same' zero padding: appending zeros after the global-max winner cannot lower any channel of the raw max;
    # the embedding of a clip equals the embedding of the same clip computed alone (already checked) and the
    # encoder is deterministic run to run
    assert torch.equal(eng.forward(x), full)


@pytest.mark.parametrize("precision", [2, 3])
@pytest.mark.parametrize("filters,emb", [(16, 32), (32, 64), (64, 128), (256, 64), (512, 512)])
def test_encoder_parity_filter_sweep(filters, emb, precision):
    """grid_search_siamese_network.py:23-25 sweeps filters in [16, 32, 64, 128] and embedding in [32..512]; 256 and
    512 are the stretch widths of SURVEY.md 8(d) C5 (block 4 then has 2048 output channels = 16 cout slabs)."""
    params = O.init_encoder_params(filters, emb, seed=filters, randomize_bn=True, random_bias=True)
    eng = _engine(filters, emb, params, precision)
    x = O.synthetic_clips(3, 6000, seed=8)
    got = eng.forward(torch.from_numpy(x[:, :, 0].copy()).cuda()).cpu().numpy()
    ref = O.encoder_forward(x, params, torch.float32)
    assert _per_clip(got, ref) <= TOL


def test_unsupported_configuration_is_an_error_not_a_fallback():
    from voicemap_b200 import _lib
    from voicemap_b200.engine import EncoderEngine
    params = O.init_encoder_params(20, 8, seed=1)   # channel counts must be multiples of 8
    eng = EncoderEngine(20, 8)
    eng.set_weights(params)
    with pytest.raises(_lib.VoicemapB200Error):
        eng.forward(torch.zeros(2, 4000, device="cuda"))


@pytest.mark.parametrize("n,t,ds", [(6, 48000, 4), (3, 16001, 4), (4, 12000, 1)])
def test_fused_preprocessing_matches_host_preprocessing(stress_params, n, t, ds):
    """predict_raw = predict(preprocess_instances(ds)(x)): decimation x[:, ::ds] and whitening with the reference's
    batch-global scale (voicemap/utils.py:22-34, 88-101) fused into block 1."""
    from voicemap_b200.models import get_baseline_convolutional_encoder
    rng = np.random.default_rng(t)
    raw = (rng.normal(0.01, 0.05, (n, t, 1)) * rng.uniform(0.3, 3.0, (n, 1, 1))).astype(np.float32)
    enc = get_baseline_convolutional_encoder(128, 64)
    enc.set_named_weights(stress_params)
    got = enc.predict_raw(raw, downsampling=ds)
    pre = O.preprocess_instances(ds)(raw.astype(np.float64))
    assert pre.shape[1] == -(-t // ds)
    ref = O.encoder_forward(pre, stress_params, torch.float32)
    assert _per_clip(got, ref) <= TOL
    assert _per_clip(got, enc.predict(pre)) <= 2e-5          # device preprocessing == host preprocessing + predict
    nowhite = enc.predict_raw(raw, downsampling=ds, whitening=False)
    assert _per_clip(nowhite, O.encoder_forward(raw[:, ::ds], stress_params, torch.float32)) <= TOL


def test_batched_n_shot_evaluation_equals_reference_loop(stress_params):
    """many tasks per launch, identical count of solved tasks for the same task sequence."""
    import pandas as pd
    from voicemap_b200 import utils
    from voicemap_b200.librispeech import LibriSpeechDataset
    from voicemap_b200.models import build_siamese_net, get_baseline_convolutional_encoder
    rng = np.random.default_rng(0)
    rows, audio = [], {}
    for spk in range(8):
        for j in range(4):
            path = f"/fake/{spk}/{j}"
            length = int(16000 * 0.3 + rng.integers(0, 2000))
            t = np.arange(length) / 16000.0
            audio[path] = 0.05 * np.sin(2 * np.pi * (120 + 35 * spk) * t + rng.uniform(0, 6)) + 0.01 * rng.normal(size=length)
            rows.append(dict(id=spk, sex="M", subset="s", minutes=1.0, name=str(spk), filepath=path, length=length,
                             seconds=length / 16000.0))
    ds = LibriSpeechDataset("s", 0.25, stochastic=False, index=pd.DataFrame(rows), reader=lambda p: (audio[p], 16000))
    enc = get_baseline_convolutional_encoder(128, 64, dropout=0.0)
    enc.set_named_weights(stress_params)
    sia = build_siamese_net(enc, (1000, 1))
    pre = utils.BatchPreProcessor("siamese", utils.preprocess_instances(4))
    for n, k, kind in ((1, 5, "siamese"), (2, 3, "siamese")):
        np.random.seed(7)
        a = utils.n_shot_task_evaluation(sia, ds, pre, 12, n, k, network_type=kind)
        np.random.seed(7)
        b = utils.n_shot_task_evaluation_batched(sia, ds, pre, 12, n, k, network_type=kind, tasks_per_launch=5)
        assert a == b


def test_reference_checkpoint_weights_fixture():
    """Real trained weights (the reference's shipped Keras checkpoint, read by voicemap_b200/keras_hdf5.py) through
    the CUDA path vs the oracle's committed outputs: filters=32, embedding 128, weighted_l1 head."""
    from voicemap_b200.models import build_siamese_net, get_baseline_convolutional_encoder
    z = np.load(os.path.join(os.path.dirname(GOLDEN), "checkpoint_f32.npz"))
    enc = get_baseline_convolutional_encoder(32, 128)
    enc.set_named_weights({k[2:]: z[k] for k in z.files if k.startswith("w_")})
    got = enc.predict(z["x"])
    assert _per_clip(got, z["emb32"]) <= TOL and _per_clip(got, z["emb64"]) <= TOL
    sia = build_siamese_net(enc, (12000, 1), "weighted_l1")
    sia.head_weights["head_kernel"] = z["head_kernel"].copy()
    sia.head_weights["head_bias"] = z["head_bias"].copy()
    prob = sia.predict([z["x"][:3], z["x"][3:]])
    assert np.abs(prob - z["prob64"]).max() < 1e-4


@pytest.mark.parametrize("n,length", [(4, 12000), (3, 1999), (2, 516), (5, 16)])
def test_first_pool_2_architecture_matches_oracle(stress_params, n, length):
    """Older reference architecture (MaxPool1D 2,2,2,2 -- the shipped checkpoint): block 1 pools 2:1 in
    two staging passes per tile.  Ragged lengths exercise the clipped second pass."""
    from voicemap_b200.models import get_baseline_convolutional_encoder
    enc = get_baseline_convolutional_encoder(128, 64, dropout=0.0, first_pool=2)
    enc.set_named_weights(stress_params)
    x = O.synthetic_clips(n, length, seed=77 + length, padded=(length == 12000))
    ref = O.encoder_forward(x, stress_params, torch.float32, pools=(2, 2, 2, 2))
    assert _per_clip(enc.predict(x), ref) <= TOL
    enc.precision = 3        # fp16 (hi, lo) planes: the merged block output is fp32-grade
    assert _per_clip(enc.predict(x), ref) <= TOL
    eng = enc._get_engine()
    hi, lo = eng.block1(torch.from_numpy(x[:, :, 0].astype(np.float32)).cuda())
    assert hi.shape == (n, length // 2, 128)
    inter = O.encoder_forward(x, stress_params, torch.float64, pools=(2, 2, 2, 2), return_intermediates=True)[1]
    got = eng.merge_planes(hi, lo).cpu().numpy()
    assert np.abs(got - inter[0]).max() <= 1e-5 * np.abs(inter[0]).max()


def test_shipped_checkpoint_in_its_own_architecture():
    """Real trained weights (checkpoint_f32.npz = the reference's shipped checkpoint) in the architecture they were
    trained in (first pool 2, embedding 128, weighted_l1 head) against the oracle."""
    from voicemap_b200.models import build_siamese_net, get_baseline_convolutional_encoder
    z = np.load(os.path.join(os.path.dirname(GOLDEN), "checkpoint_f32.npz"))
    w = {k[2:]: z[k] for k in z.files if k.startswith("w_")}
    enc = get_baseline_convolutional_encoder(32, 128, first_pool=2)
    enc.set_named_weights(w)
    ref64 = O.encoder_forward(z["x"], w, torch.float64, pools=(2, 2, 2, 2))
    assert _per_clip(enc.predict(z["x"]), ref64) <= TOL
    sia = build_siamese_net(enc, (12000, 1), "weighted_l1")
    sia.head_weights["head_kernel"] = z["head_kernel"].copy()
    sia.head_weights["head_bias"] = z["head_bias"].copy()
    prob = sia.predict([z["x"][:3], z["x"][3:]])
    refp, _ = O.siamese_head(ref64[:3], ref64[3:], z["head_kernel"].reshape(-1).astype(np.float64),
                             float(z["head_bias"][0]), "weighted_l1")
    assert np.abs(prob - refp).max() < 1e-4


@pytest.mark.parametrize("precision", [2, 3])
def test_large_batch_of_short_clips(stress_params, precision):
    """n_seconds sweep corner (1 s clips, L = 4000) at a batch well beyond one wave of tiles per SM."""
    eng = _engine(128, 64, stress_params, precision)
    g = torch.Generator().manual_seed(9)
    x = (O.WHITEN_RMS * torch.randn(1536, 4000, generator=g)).cuda()
    full = eng.forward(x).clone()
    assert torch.isfinite(full).all()
    idx = [0, 777, 1535]
    ref = O.encoder_forward(x[idx].cpu().numpy()[:, :, None], stress_params, torch.float32)
    assert _per_clip(full[idx].cpu().numpy(), ref) <= TOL
    assert torch.equal(eng.forward(x[700:800].contiguous()), full[700:800])


@pytest.mark.parametrize("distance", ["euclidean", "cosine", "dot_product"])
@pytest.mark.parametrize("k,n,width", [(5, 1, 64), (5, 5, 64), (20, 5, 128), (3, 2, 33)])
def test_nshot_scoring_on_device_equals_reference_rule(distance, k, n, width):
    """vm_nshot_score (class means, distance, arg-min on the device) against the numpy statement of
    voicemap/utils.py:156-212 (utils._SCORES), incl. an exact tie (first minimum wins, as np.argmin)."""
    from voicemap_b200 import utils
    rng = np.random.default_rng(k * 100 + n)
    tasks = 37
    query = rng.normal(size=(tasks, width)).astype(np.float32)
    support = rng.normal(size=(tasks, k * n, width)).astype(np.float32)
    support[0, n:2 * n] = support[0, :n]                      # task 0: classes 0 and 1 coincide and are nearest:
    query[0] = support[0, :n].mean(axis=0)                    # an exact tie that the first class must win
    best = utils.nshot_best_class_device(torch.from_numpy(query).cuda(),
                                         torch.from_numpy(support.reshape(tasks * k * n, width)).cuda(), k, n, distance)
    want = [int(np.argmin(utils._SCORES[distance](query[t].astype(np.float64), support[t].astype(np.float64), k, n)))
            for t in range(tasks)]
    assert best.cpu().tolist() == want
    if distance != "dot_product":     # (a longer vector elsewhere can out-project the coinciding pair)
        assert want[0] == 0


@pytest.mark.parametrize("loss", ["contrastive_loss", "binary_crossentropy"])
def test_siamese_probability_and_loss_at_config3_shape(stress_params, loss):
    """BASELINE config[2]'s shape in eval mode: 128 pairs x 12000, filters 128 -- CUDA encoder (both branches as one
    256-clip launch) -> fused head + loss kernel, against the fp64 oracle at |d| / |ref| <= 1e-4 for the probabilities
    and the loss (voicemap/models.py:64-69, voicemap/utils.py:77-85)."""
    from voicemap_b200.keras_compat import Adam
    from voicemap_b200 import models as M
    from voicemap_b200 import utils
    pairs, length = 128, 12000
    enc = M.get_baseline_convolutional_encoder(128, 64, dropout=0.0)
    enc.set_named_weights(stress_params)
    sia = M.build_siamese_net(enc, (length, 1))
    x = O.synthetic_clips(2 * pairs, length, seed=77)
    y = (np.arange(pairs) >= pairs // 2).astype(np.float64)[:, None]
    ref_emb = O.encoder_forward(x, stress_params, torch.float64)
    spread = float(np.median(np.linalg.norm(ref_emb[:pairs] - ref_emb[pairs:], axis=1)))
    w, b = np.array([[2.0 / spread]], np.float32), np.array([-1.5], np.float32)     # un-saturated sigmoid
    sia.set_weights(enc.get_weights() + [w, b])
    ref_prob, _ = O.siamese_head(ref_emb[:pairs], ref_emb[pairs:], float(w[0, 0]), float(b[0]))
    assert 0.02 < ref_prob.min() and ref_prob.max() < 0.98
    prob = sia.predict([x[:pairs], x[pairs:]])
    assert np.abs(prob - ref_prob).max() <= 1e-4 * np.abs(ref_prob).max()
    sia.compile(loss=utils.contrastive_loss if loss == "contrastive_loss" else loss, optimizer=Adam())
    got = sia.test_on_batch([x[:pairs], x[pairs:]], y)
    want = O.contrastive_loss(y, ref_prob) if loss == "contrastive_loss" else O.binary_crossentropy(y, ref_prob)
    assert abs(got - float(want)) <= 1e-4 * abs(float(want)), (got, float(want))
