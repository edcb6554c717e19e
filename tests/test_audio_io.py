"""Audio decoding for the batcher (voicemap_b200/audio_io.py over libvoicemap_io.so), CPU only.

Anchors: two complete FLAC streams from the appendix of RFC 9639 ("decoding examples" 1 and 3), whose STREAMINFO
blocks carry the MD5 of the original PCM -- the decoder's output must hash to it.  Everything else round-trips streams
written by tests/flac_writer.py (an independent pure-Python encoder) through the C decoder, bit-exactly.
"""
import ctypes
import hashlib
import os
import re
import wave

import numpy as np
import pytest

from flac_writer import encode_flac, id3v2_tag
from voicemap_b200 import audio_io
from voicemap_b200.build import IO_LIB_PATH, build_io_library

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# RFC 9639 appendix D.1: one stereo 16-bit sample pair, two VERBATIM subframes with wasted bits
RFC_EXAMPLE_1 = bytes.fromhex(
    "664c6143800000221000100000000f00000f0ac442f0000000013e84b41807dc690307586a3dad1a2e0f"
    "fff869180000bf0358fd03128baa9a")
# RFC 9639 appendix D.3: 24 mono 8-bit samples at 32 kHz, one LPC subframe (order 3, 4-bit coefficients, Rice residual)
RFC_EXAMPLE_3 = bytes.fromhex(
    "664c61438000002210001000" "00001f00001f07d000700000" "0018f8f9e396f5cbcfc6dc80" "7f9977906b32fff868020017"
    "e944004f6f313d1047d227cb" "6d090831452bdc2822228057" "a3")


@pytest.fixture(scope="module", autouse=True)
def io_library():
    build_io_library()
    return IO_LIB_PATH


def _speechlike(n, channels=1, bps=16, seed=0):
    """Smooth, predictable signal + noise, full scale for `bps` (so LPC / fixed predictors have something to do)."""
    rng = np.random.default_rng(seed)
    t = np.arange(n)[:, None]
    freq = rng.uniform(0.003, 0.05, size=(1, channels))
    x = 0.6 * np.sin(2 * np.pi * freq * t + rng.uniform(0, 6, (1, channels))) + 0.02 * rng.standard_normal((n, channels))
    if channels == 2:
        x[:, 1] = 0.8 * x[:, 0] + 0.1 * x[:, 1]  # correlated channels, as stereo decorrelation expects
    full = (1 << (bps - 1)) - 1
    return np.clip(np.round(x * full), -full - 1, full).astype(np.int64)


def _decode_int(stream):
    pcm, rate = audio_io.decode_flac(stream, dtype="int32")
    return pcm, rate


# ----------------------------------------------------------------------------------------- C ABI surface
def test_header_symbols_exported_and_bound(io_library):
    text = open(os.path.join(ROOT, "include", "voicemap_io.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    declared = sorted(set(re.findall(r"\b(vmio_[a-z0-9_]+)\s*\(", text)))
    assert declared == sorted(audio_io.SIGNATURES) and len(declared) == 8
    lib = ctypes.CDLL(io_library)
    for name in declared:
        assert hasattr(lib, name)
    assert audio_io.load().vmio_version() == 100


def test_struct_layout_matches_header():
    assert ctypes.sizeof(audio_io.FlacInfo) == 48  # 5 x u32, pad, u64, 16 bytes
    assert audio_io.FlacInfo.total_samples.offset == 24 and audio_io.FlacInfo.md5.offset == 32


# ----------------------------------------------------------------------------------------- known answers
def test_rfc9639_example_1_verbatim_wasted_bits():
    info = audio_io.flac_info(RFC_EXAMPLE_1)
    assert (info["samplerate"], info["channels"], info["bits_per_sample"], info["frames"]) == (44100, 2, 16, 1)
    pcm, rate = _decode_int(RFC_EXAMPLE_1)
    assert rate == 44100 and pcm.tolist() == [[25588, 10416]]
    assert hashlib.md5(pcm.astype("<i2").tobytes()).digest() == info["md5"]
    assert info["md5"].hex() == "3e84b41807dc690307586a3dad1a2e0f"


def test_rfc9639_example_3_lpc():
    info = audio_io.flac_info(RFC_EXAMPLE_3)
    assert (info["samplerate"], info["channels"], info["bits_per_sample"], info["frames"]) == (32000, 1, 8, 24)
    pcm, _ = _decode_int(RFC_EXAMPLE_3)
    assert pcm.tolist() == [0, 79, 111, 78, 8, -61, -90, -68, -13, 42, 67, 53, 13, -27, -46, -38, -12, 14, 24, 19, 6,
                            -4, -5, 0]
    assert hashlib.md5(pcm.astype("<i1").tobytes()).digest() == info["md5"]
    samples, _ = audio_io.decode_flac(RFC_EXAMPLE_3)           # soundfile convention: int / 2^(bits-1), float64
    assert samples.dtype == np.float64 and samples.shape == (24,)
    np.testing.assert_array_equal(samples, pcm / 128.0)


# ----------------------------------------------------------------------------------------- round trips
@pytest.mark.parametrize("kind,options", [
    ("verbatim", {}),
    ("fixed", {"fixed_order": 0}),
    ("fixed", {"fixed_order": 1, "partition_order": 2}),
    ("fixed", {"fixed_order": 2, "partition_order": 4}),
    ("fixed", {"fixed_order": 3, "method": 1, "partition_order": 3}),
    ("fixed", {"fixed_order": 4, "partition_order": 1, "escape_partitions": (1,)}),
    ("lpc", {"lpc_order": 1, "lpc_precision": 5}),
    ("lpc", {"lpc_order": 8, "lpc_precision": 12, "partition_order": 3}),
    ("lpc", {"lpc_order": 12, "lpc_precision": 15, "method": 1, "partition_order": 2, "escape_partitions": (0, 3)}),
    ("lpc", {"lpc_order": 32, "lpc_precision": 14, "partition_order": 0}),
])
def test_subframe_types_round_trip_16bit_mono(kind, options):
    """LibriSpeech's shape: 16 kHz, 16 bit, mono, blocks of 4096 with a ragged last block."""
    pcm = _speechlike(4096 + 1000, seed=3)[:, 0]
    got, rate = _decode_int(encode_flac(pcm[:4096], 16000, 16, 4096, kind=kind, **options))
    assert rate == 16000
    np.testing.assert_array_equal(got, pcm[:4096])
    # 1000 samples do not split into 2^k partitions beyond k = 3: the ragged stream uses a single partition
    ragged = {**options, "partition_order": 0, "escape_partitions": ()}
    got, _ = _decode_int(encode_flac(pcm, 16000, 16, 4096, kind=kind, **ragged))
    np.testing.assert_array_equal(got, pcm)


@pytest.mark.parametrize("stereo", ["independent", "left_side", "side_right", "mid_side"])
@pytest.mark.parametrize("bps", [8, 16, 24])
def test_stereo_decorrelation_and_bit_depths(stereo, bps):
    pcm = _speechlike(1152 * 2, channels=2, bps=bps, seed=bps)
    stream = encode_flac(pcm, 44100, bps, 1152, kind="lpc", lpc_order=6, lpc_precision=13, partition_order=2,
                         stereo=stereo)
    got, rate = _decode_int(stream)
    assert rate == 44100 and got.shape == (2304, 2)
    np.testing.assert_array_equal(got, pcm)
    info = audio_io.flac_info(stream)
    width = bps // 8
    raw = b"".join(int(v).to_bytes(width, "little", signed=True) for v in got.reshape(-1))
    assert hashlib.md5(raw).digest() == info["md5"]
    samples, _ = audio_io.decode_flac(stream)
    np.testing.assert_array_equal(samples, pcm / float(1 << (bps - 1)))
    assert -1.0 <= samples.min() and samples.max() < 1.0


def test_extreme_sample_values_and_32bit():
    """Full-scale alternation defeats every predictor (residuals exceed the sample width); 32-bit side channel = 33 bits."""
    for bps in (16, 32):
        lo, hi = -(1 << (bps - 1)), (1 << (bps - 1)) - 1
        pcm = np.array([lo, hi] * 96, dtype=np.int64)
        for kind, opts in (("verbatim", {}), ("fixed", {"fixed_order": 2 if bps == 16 else 1, "method": 1})):
            got, _ = _decode_int(encode_flac(pcm, 16000, bps, 192, kind=kind, **opts))
            np.testing.assert_array_equal(got, pcm)
    stereo = np.stack([np.array([(1 << 31) - 1, -(1 << 31)] * 96), np.array([-(1 << 31), (1 << 31) - 1] * 96)], axis=1)
    for mode in ("left_side", "side_right", "mid_side"):
        got, _ = _decode_int(encode_flac(stereo, 48000, 32, 192, kind="verbatim", stereo=mode))
        np.testing.assert_array_equal(got, stereo)


def test_constant_and_wasted_bits_and_silence():
    silence = np.zeros(512, dtype=np.int64)
    got, _ = _decode_int(encode_flac(silence, 16000, 16, 256, kind="auto"))
    np.testing.assert_array_equal(got, silence)
    dc = np.full(512, -1234, dtype=np.int64)
    got, _ = _decode_int(encode_flac(dc, 16000, 16, 256, kind="auto"))
    np.testing.assert_array_equal(got, dc)
    coarse = _speechlike(1024, seed=9)[:, 0] // 64 * 64          # six wasted bits
    for kind in ("verbatim", "fixed", "lpc"):
        got, _ = _decode_int(encode_flac(coarse, 16000, 16, 512, kind=kind))
        np.testing.assert_array_equal(got, coarse)
    zero_residual = np.arange(256, dtype=np.int64) * 3             # order-2 residual is all zero: escape with 0 raw bits
    got, _ = _decode_int(encode_flac(zero_residual, 16000, 16, 256, kind="fixed", fixed_order=2,
                                     escape_partitions=(0,)))
    np.testing.assert_array_equal(got, zero_residual)


@pytest.mark.parametrize("rate", [16000, 8000, 96000, 11000, 11025, 380000])
def test_header_forms_rates(rate):
    pcm = _speechlike(600, seed=1)[:, 0]
    got, got_rate = _decode_int(encode_flac(pcm, rate, 16, 200, kind="fixed"))   # 200: 8-bit block size field
    assert got_rate == rate
    np.testing.assert_array_equal(got, pcm)
    got, got_rate = _decode_int(encode_flac(pcm, rate, 16, 300, kind="fixed", rate_from_streaminfo=True,
                                            bits_from_streaminfo=True))           # 300: 16-bit block size field
    assert got_rate == rate
    np.testing.assert_array_equal(got, pcm)


def test_variable_blocksize_unknown_length_and_tags():
    pcm = _speechlike(192 + 576 + 1000 + 16, seed=5)[:, 0]
    padding = (1, bytes(100))                                       # a PADDING block after STREAMINFO
    stream = encode_flac(pcm, 16000, 16, kind="fixed", blocksizes=[192, 576, 1000, 16], record_length=False,
                         record_md5=False, extra_metadata=[padding], id3v2=id3v2_tag(300), id3v1=True)
    info = audio_io.flac_info(stream)
    assert info["frames"] == 0 and info["md5"] == bytes(16) and (info["min_blocksize"], info["max_blocksize"]) == (16, 1000)
    got, _ = _decode_int(stream)
    np.testing.assert_array_equal(got, pcm)


def test_many_frames_frame_numbers_cross_utf8_widths():
    pcm = _speechlike(16 * 2100, seed=6)[:, 0]                      # 2100 frames: numbers need 1, 2 and 3 bytes
    got, _ = _decode_int(encode_flac(pcm, 16000, 16, 16, kind="fixed", fixed_order=1))
    np.testing.assert_array_equal(got, pcm)
    big = encode_flac(pcm[:64], 16000, 16, kind="verbatim", blocksizes=[16] * 4)
    got, _ = _decode_int(big)
    np.testing.assert_array_equal(got, pcm[:64])


# ----------------------------------------------------------------------------------------- partial decoding
@pytest.mark.parametrize("layout", ["fixed_blocks", "variable_blocks", "unknown_length", "stereo"])
def test_range_decode_equals_slice_of_full_decode(layout):
    rng = np.random.default_rng(12)
    if layout == "stereo":
        pcm = _speechlike(256 * 40 + 77, channels=2, seed=4)
        stream = encode_flac(pcm, 16000, 16, 256, kind="fixed", fixed_order=2, stereo="mid_side")
    else:
        pcm = _speechlike(256 * 40 + 77, seed=4)[:, 0]
        if layout == "variable_blocks":
            sizes = [192, 576, 256, 1000] * 5 + [256 * 40 + 77 - 5 * 2024]
            stream = encode_flac(pcm, 16000, 16, kind="lpc", lpc_order=4, blocksizes=sizes)
        else:
            stream = encode_flac(pcm, 16000, 16, 256, kind="lpc", lpc_order=4, partition_order=0,
                                 record_length=(layout != "unknown_length"))
    total = len(pcm)
    spans = [(0, total), (0, 1), (total - 1, 1), (255, 2), (256, 256), (5000, 3000), (total - 100, 500), (total, 10),
             (total + 5000, 10), (123, 0)]
    spans += [(int(a), int(b)) for a, b in zip(rng.integers(0, total, 25), rng.integers(1, 4000, 25))]
    for start, count in spans:
        got, rate = audio_io.decode_flac_range(stream, start, count, dtype="int32")
        assert rate == 16000
        np.testing.assert_array_equal(got, pcm[start:start + count], err_msg=f"span {start}+{count}")
    samples, _ = audio_io.decode_flac_range(stream, 1000, 48)
    np.testing.assert_array_equal(samples, pcm[1000:1048] / 32768.0)


def test_range_decode_survives_sync_codes_inside_audio():
    """VERBATIM audio made of 0xFFF8 words looks like a frame header at every byte pair: resynchronisation must reject
    the impostors (header CRC-8 / frame CRC-16) and still land on real frames."""
    pcm = np.full(4096 * 3, -8, dtype=np.int64)          # 0xfff8 as a 16-bit sample
    pcm[::7] = 0x18c                                      # break the constant so the subframes stay verbatim
    stream = encode_flac(pcm, 16000, 16, 1024, kind="verbatim", allow_wasted=False)
    assert stream.count(b"\xff\xf8") > 5000
    for start in (0, 1500, 5000, 9000, 12000):
        got, _ = audio_io.decode_flac_range(stream, start, 700, dtype="int32")
        np.testing.assert_array_equal(got, pcm[start:start + 700])


# ----------------------------------------------------------------------------------------- damage is detected
def test_corruption_is_an_error_not_garbage():
    pcm = _speechlike(2048, seed=7)[:, 0]
    stream = bytearray(encode_flac(pcm, 16000, 16, 1024, kind="lpc", partition_order=2))
    first_frame = stream.index(b"\xff\xf8", 42)
    cases = {
        "CRC-8": first_frame + 2,          # block size / rate byte of the first frame header
        "CRC-16": len(stream) - 40,        # inside the second frame's residual
    }
    for what, where in cases.items():
        damaged = bytearray(stream)
        damaged[where] ^= 0x10
        with pytest.raises(audio_io.AudioDecodeError) as err:
            audio_io.decode_flac(bytes(damaged))
        assert what in str(err.value) or "header" in str(err.value) or "residual" in str(err.value), str(err.value)
    with pytest.raises(audio_io.AudioDecodeError, match="truncated"):
        audio_io.decode_flac(bytes(stream[:len(stream) - 7]))
    with pytest.raises(audio_io.AudioDecodeError, match="not a FLAC"):
        audio_io.decode_flac(b"RIFF" + bytes(100))
    with pytest.raises(audio_io.AudioDecodeError, match="synchronisation"):
        audio_io.decode_flac(bytes(stream[:first_frame]) + b"\x00" * 32)
    with pytest.raises(audio_io.AudioDecodeError):
        audio_io.decode_flac(b"")


def test_capacity_is_checked(io_library):
    """A stream that holds more audio than its STREAMINFO promises must not overrun the caller's buffer."""
    pcm = _speechlike(1024, seed=8)[:, 0]
    stream = np.frombuffer(encode_flac(pcm, 16000, 16, 256, kind="fixed", record_length=False), dtype=np.uint8)
    lib = audio_io.load()
    out = np.zeros(512 + 8, dtype=np.float64)
    rc = lib.vmio_flac_decode(stream.ctypes.data, stream.size, None, out.ctypes.data, 512, None)
    assert rc == -9 and not out[512:].any()
    assert lib.vmio_flac_decode(stream.ctypes.data, stream.size, None, None, 0, None) == 1024


# ----------------------------------------------------------------------------------------- files
def test_read_dispatch_flac_wav_and_unknown(tmp_path):
    pcm = _speechlike(5000, seed=11)[:, 0]
    flac_path = tmp_path / "utt.flac"
    flac_path.write_bytes(encode_flac(pcm, 16000, 16, 4096, kind="lpc", partition_order=0))
    samples, rate = audio_io.read(str(flac_path))
    assert rate == 16000
    np.testing.assert_array_equal(samples, pcm / 32768.0)
    assert audio_io.flac_info(str(flac_path))["frames"] == 5000

    # the C file entry point gives the same samples as the in-memory one
    lib = audio_io.load()
    out = np.empty(5000, dtype=np.int32)
    info = audio_io.FlacInfo()
    assert lib.vmio_flac_read_file(os.fsencode(str(flac_path)), out.ctypes.data, None, 5000, ctypes.byref(info)) == 5000
    np.testing.assert_array_equal(out, pcm)
    assert info.sample_rate == 16000 and info.total_samples == 5000
    assert lib.vmio_flac_read_file(os.fsencode(str(tmp_path / "missing.flac")), None, None, 0, None) == -12

    wav_path = tmp_path / "utt.wav"
    with wave.open(str(wav_path), "wb") as handle:
        handle.setnchannels(1)
        handle.setsampwidth(2)
        handle.setframerate(16000)
        handle.writeframes(pcm.astype("<i2").tobytes())
    samples, rate = audio_io.read(str(wav_path))
    assert rate == 16000
    np.testing.assert_array_equal(samples, pcm / 32768.0)

    with pytest.raises((audio_io.AudioDecodeError, RuntimeError)):
        audio_io.read(str(tmp_path / "utt.ogg"))
    with pytest.raises(audio_io.AudioDecodeError):
        audio_io.read(str(tmp_path / "missing.flac"))


def test_wav_widths(tmp_path):
    rng = np.random.default_rng(0)
    for width in (1, 2, 3, 4):
        bits = 8 * width
        pcm = rng.integers(-(1 << (bits - 1)), 1 << (bits - 1), size=(300, 2), dtype=np.int64)
        if width == 1:
            raw = (pcm + 128).astype(np.uint8).tobytes()
        else:
            raw = b"".join(int(v).to_bytes(width, "little", signed=True) for v in pcm.reshape(-1))
        path = tmp_path / f"w{width}.wav"
        with wave.open(str(path), "wb") as handle:
            handle.setnchannels(2)
            handle.setsampwidth(width)
            handle.setframerate(8000)
            handle.writeframes(raw)
        samples, rate = audio_io.read_wav(path)
        assert rate == 8000 and samples.shape == (300, 2)
        np.testing.assert_array_equal(samples, pcm / float(1 << (bits - 1)))


def test_read_many_keeps_order_and_matches_serial(tmp_path):
    paths = []
    for i in range(12):
        pcm = _speechlike(3000 + 17 * i, seed=20 + i)[:, 0]
        path = tmp_path / f"{i}.flac"
        path.write_bytes(encode_flac(pcm, 16000, 16, 1024, kind="fixed", fixed_order=2))
        paths.append(str(path))
    serial = [audio_io.read(p) for p in paths]
    threaded = audio_io.read_many(paths, workers=4)
    assert [len(s) for s, _ in threaded] == [3000 + 17 * i for i in range(12)]
    for (a, ra), (b, rb) in zip(serial, threaded):
        assert ra == rb
        np.testing.assert_array_equal(a, b)


def test_batcher_reads_a_real_flac_corpus(tmp_path):
    """The LibriSpeech batcher on genuine FLAC files with its default reader: lengths come from the stream headers,
    clips are the decoded PCM / 32768 (what soundfile.read returns at voicemap/librispeech.py:104), and threaded
    decoding consumes the random stream exactly like the serial loop."""
    from voicemap_b200.librispeech import LibriSpeechDataset
    root = tmp_path / "data" / "LibriSpeech"
    pcm_of = {}
    for spk in (14, 16, 21):
        for u in range(3):
            d = root / "dev-clean" / str(spk) / "100"
            d.mkdir(parents=True, exist_ok=True)
            pcm = _speechlike(16000 + 4000 * u + 100 * spk, seed=spk * 10 + u)[:, 0]
            path = d / f"{spk}-100-{u:04d}.flac"
            path.write_bytes(encode_flac(pcm, 16000, 16, 4096, kind="lpc", lpc_order=8, partition_order=0))
            pcm_of[str(path)] = pcm
    (root / "SPEAKERS.TXT").write_text("; c\n14 | F | dev-clean | 25.0 | A\n16 | M | dev-clean | 25.0 | B\n"
                                       "21 | M | dev-clean | 25.0 | C\n")
    ds = LibriSpeechDataset("dev-clean", 0.5, stochastic=False, data_path=str(tmp_path), cache=False)
    assert len(ds) == 9 and ds.num_classes() == 3
    assert sorted(ds.df["length"]) == sorted(len(v) for v in pcm_of.values())
    for i in range(len(ds)):
        clip, label = ds[i]
        path = ds.df["filepath"][i]
        assert clip.dtype == np.float64 and label == int(os.path.basename(path).split("-")[0])
        np.testing.assert_array_equal(clip, pcm_of[path][:8000] / 32768.0)

    # stochastic items: the fragment decoded alone equals the reference recipe on the fully decoded file with the same
    # random stream (voicemap/librispeech.py:105-124), padding included
    padded = LibriSpeechDataset("dev-clean", 1.6, stochastic=True, pad=True, data_path=str(tmp_path), cache=False)
    assert len(padded) == 9
    for i in range(len(padded)):
        np.random.seed(100 + i)
        clip, _ = padded[i]
        np.random.seed(100 + i)
        want = padded._fragment(pcm_of[padded.df["filepath"][i]] / 32768.0)
        assert clip.shape == (25600,)
        np.testing.assert_array_equal(clip, want)

    batches = []
    for workers in (1, 4):
        stochastic = LibriSpeechDataset("dev-clean", 0.5, stochastic=True, data_path=str(tmp_path), cache=False,
                                        decode_workers=workers)
        np.random.seed(5)
        (left, right), labels = stochastic.build_verification_batch(4)
        (query, qlabel), (support, slabels) = stochastic.build_n_shot_task(2, 1)
        assert left.shape == right.shape == (4, 8000, 1) and labels.ravel().tolist() == [0, 0, 1, 1]
        assert support.shape == (2, 8000) and slabels[0] == qlabel
        batches.append((left, right, query, support))
    for a, b in zip(*batches):
        np.testing.assert_array_equal(a, b)


def test_read_fragments_batch_call(tmp_path):
    """The native batch reader: crop, leading / trailing zeros, thread counts, and its error reports."""
    paths, pcms = [], []
    for i in range(9):
        pcm = _speechlike(5000 + 300 * i, seed=40 + i)[:, 0]
        path = tmp_path / f"{i}.flac"
        path.write_bytes(encode_flac(pcm, 16000, 16, 1024, kind="fixed", fixed_order=2))
        paths.append(str(path))
        pcms.append(pcm)
    rng = np.random.default_rng(3)
    want = 4000
    starts = [int(rng.integers(0, len(p) - 3000)) for p in pcms]
    counts = [int(rng.integers(0, 3000)) for _ in pcms]
    counts[0], counts[1] = 0, want                 # an empty fragment and one that fills the clip
    starts[1] = 17
    leads = [int(rng.integers(0, want - c + 1)) for c in counts]
    expect = np.zeros((9, want))
    for i, pcm in enumerate(pcms):
        expect[i, leads[i]:leads[i] + counts[i]] = pcm[starts[i]:starts[i] + counts[i]] / 32768.0
    for workers in (1, 3, 16):
        np.testing.assert_array_equal(audio_io.read_fragments(paths, starts, counts, want, leads, workers=workers), expect)
    no_lead = audio_io.read_fragments(paths[:2], [5, 6], [10, 20], 25)
    np.testing.assert_array_equal(no_lead[1, :20], pcms[1][6:26] / 32768.0)
    assert not no_lead[0, 10:].any() and audio_io.read_fragments([], [], [], 10).shape == (0, 10)

    with pytest.raises(audio_io.AudioDecodeError, match="3.flac.*fewer samples"):     # row 3 asks past its file's end
        audio_io.read_fragments(paths, [0, 0, 0, len(pcms[3]) - 50, 0, 0, 0, 0, 0], [100] * 9, want, [0] * 9, workers=2)
    with pytest.raises(audio_io.AudioDecodeError, match="could not be read"):
        audio_io.read_fragments(paths[:2] + [str(tmp_path / "nope.flac")], [0] * 3, [10] * 3, want)
    with pytest.raises(audio_io.AudioDecodeError, match="exceeds"):
        audio_io.read_fragments(paths[:1], [0], [want], want, [1])
    stereo = tmp_path / "stereo.flac"
    stereo.write_bytes(encode_flac(_speechlike(2048, channels=2), 16000, 16, 1024, kind="fixed"))
    with pytest.raises(audio_io.AudioDecodeError, match="unsupported"):
        audio_io.read_fragments([str(stereo)], [0], [10], want)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(audio_io, "_lib", None)
    monkeypatch.setattr(audio_io, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(audio_io.AudioDecodeError, match="voicemap_b200.build"):
        audio_io.load()
