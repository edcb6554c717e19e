"""Audio decoding for the batcher: ``read(path) -> (float64 samples, sample rate)`` with the result convention of
``soundfile.read``, which is what the reference calls on every utterance (voicemap/librispeech.py:104,267).

* ``.flac`` -- decoded by ``libvoicemap_io.so`` (our own C decoder, ``csrc/vm_flac.c`` behind
  ``include/voicemap_io.h``; ctypes releases the GIL, so ``read_many`` decodes a batch on a thread pool);
* ``.wav``  -- the standard library's ``wave`` module (integer PCM, 8/16/24/32 bit);
* anything else -- ``soundfile`` if it is installed, otherwise an error that says so.

``flac_info(path)`` returns the stream header without decoding (length, rate, channels), which turns corpus indexing
from "decode every file" (voicemap/librispeech.py:267-275) into "read 4 KB of every file".
"""
from __future__ import annotations

import ctypes as C
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvoicemap_io.so")


class AudioDecodeError(RuntimeError):
    pass


class FlacInfo(C.Structure):
    _fields_ = [("sample_rate", C.c_uint32), ("channels", C.c_uint32), ("bits_per_sample", C.c_uint32),
                ("min_blocksize", C.c_uint32), ("max_blocksize", C.c_uint32), ("total_samples", C.c_uint64),
                ("md5", C.c_uint8 * 16)]


_u8p, _i32p, _f64p, _infop = C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(FlacInfo)
# name -> (restype, argtypes); tests check this table against include/voicemap_io.h
SIGNATURES = {
    "vmio_version": (C.c_int, []),
    "vmio_error_string": (C.c_char_p, [C.c_int]),
    "vmio_flac_probe": (C.c_int, [_u8p, C.c_size_t, _infop]),
    "vmio_flac_decode": (C.c_int64, [_u8p, C.c_size_t, _i32p, _f64p, C.c_uint64, _infop]),
    "vmio_flac_decode_range": (C.c_int64, [_u8p, C.c_size_t, C.c_uint64, C.c_uint64, _i32p, _f64p, _infop]),
    "vmio_flac_read_fragments": (C.c_int, [C.POINTER(C.c_char_p), C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                           C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_int64)]),
    "vmio_flac_read_file": (C.c_int64, [C.c_char_p, _i32p, _f64p, C.c_uint64, _infop]),
    "vmio_flac_probe_file": (C.c_int, [C.c_char_p, _infop]),
}

_lib = None
DEFAULT_WORKERS = None   # decoder threads when a call does not say; None = min(items, cores, 16)


def load():
    """Load libvoicemap_io.so once.  Raises if it has not been built (``python -m voicemap_b200.build``)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise AudioDecodeError(f"{LIB_PATH} not found: build it with `python -m voicemap_b200.build`")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def _fail(code, what):
    raise AudioDecodeError(f"{what}: {load().vmio_error_string(int(code)).decode()} (code {int(code)})")


def _as_bytes_array(data):
    arr = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    if arr.dtype != np.uint8 or arr.ndim != 1:
        raise AudioDecodeError("FLAC data must be bytes or a 1-D uint8 array")
    return np.ascontiguousarray(arr)


def flac_info(source):
    """Stream header of a FLAC file (path) or stream (bytes / uint8 array) as a dict; nothing is decoded."""
    lib = load()
    info = FlacInfo()
    if isinstance(source, (str, os.PathLike)):
        rc = lib.vmio_flac_probe_file(os.fsencode(source), C.byref(info))
    else:
        arr = _as_bytes_array(source)
        rc = lib.vmio_flac_probe(arr.ctypes.data, arr.size, C.byref(info))
    if rc != 0:
        _fail(rc, f"flac_info({source if isinstance(source, (str, os.PathLike)) else '<memory>'})")
    return {"samplerate": int(info.sample_rate), "channels": int(info.channels),
            "bits_per_sample": int(info.bits_per_sample), "frames": int(info.total_samples),
            "min_blocksize": int(info.min_blocksize), "max_blocksize": int(info.max_blocksize),
            "md5": bytes(info.md5)}


def decode_flac(data, dtype="float64"):
    """Decode a FLAC stream held in memory.  Returns ``(samples, rate)``; samples have shape (frames,) for mono and
    (frames, channels) otherwise -- float64 in [-1, 1) (``soundfile.read``'s default) or the raw int32 PCM."""
    if dtype not in ("float64", "int32"):
        raise AudioDecodeError("dtype must be 'float64' or 'int32'")
    lib = load()
    arr = _as_bytes_array(data)
    info = FlacInfo()
    rc = lib.vmio_flac_probe(arr.ctypes.data, arr.size, C.byref(info))
    if rc != 0:
        _fail(rc, "decode_flac")
    frames = int(info.total_samples)
    if frames == 0:  # length not recorded: a first pass counts (and checks) the frames
        frames = lib.vmio_flac_decode(arr.ctypes.data, arr.size, None, None, 0, None)
        if frames < 0:
            _fail(frames, "decode_flac")
    out = np.empty((frames, int(info.channels)), dtype=dtype)
    ptrs = (out.ctypes.data, None) if dtype == "int32" else (None, out.ctypes.data)
    got = lib.vmio_flac_decode(arr.ctypes.data, arr.size, ptrs[0], ptrs[1], frames, None)
    if got < 0:
        _fail(got, "decode_flac")
    out = out[:got]
    return (out[:, 0] if info.channels == 1 else out), int(info.sample_rate)


def decode_flac_range(data, start, count, dtype="float64"):
    """Samples [start, start + count) of a FLAC stream held in memory (fewer if the stream ends first); only the frames
    that overlap the range are decoded.  Same conventions as ``decode_flac``."""
    if dtype not in ("float64", "int32"):
        raise AudioDecodeError("dtype must be 'float64' or 'int32'")
    if start < 0 or count < 0:
        raise AudioDecodeError("start and count must be non-negative")
    lib = load()
    arr = _as_bytes_array(data)
    info = FlacInfo()
    rc = lib.vmio_flac_probe(arr.ctypes.data, arr.size, C.byref(info))
    if rc != 0:
        _fail(rc, "decode_flac_range")
    out = np.empty((count, int(info.channels)), dtype=dtype)
    ptrs = (out.ctypes.data, None) if dtype == "int32" else (None, out.ctypes.data)
    got = lib.vmio_flac_decode_range(arr.ctypes.data, arr.size, start, count, ptrs[0], ptrs[1], None) if count else 0
    if got < 0:
        _fail(got, "decode_flac_range")
    out = out[:got]
    return (out[:, 0] if info.channels == 1 else out), int(info.sample_rate)


def read_flac_range(path, start, count, dtype="float64"):
    try:
        data = np.fromfile(path, dtype=np.uint8)
    except OSError as exc:
        raise AudioDecodeError(f"cannot read {path}: {exc}") from exc
    try:
        return decode_flac_range(data, start, count, dtype)
    except AudioDecodeError as exc:
        raise AudioDecodeError(f"{path}: {exc}") from None


def read_fragments(paths, starts, counts, want, leads=None, workers=None):
    """One batch of clips in one native call: row i of the returned float64 (n, want) array is ``leads[i]`` zeros,
    samples [starts[i], starts[i] + counts[i]) of the mono FLAC file ``paths[i]`` and zeros up to ``want``.  Files are
    read and decoded (only the frames under each fragment) on ``workers`` threads inside the library."""
    lib = load()
    n = len(paths)
    out = np.empty((n, int(want)), dtype=np.float64)
    if n == 0:
        return out
    if workers is None:
        workers = DEFAULT_WORKERS if DEFAULT_WORKERS is not None else min(n, os.cpu_count() or 1, 16)
    encoded = [os.fsencode(p) for p in paths]
    c_paths = (C.c_char_p * n)(*encoded)
    c_start = np.ascontiguousarray(starts, dtype=np.uint64)
    c_count = np.ascontiguousarray(counts, dtype=np.uint64)
    c_lead = np.ascontiguousarray(leads if leads is not None else np.zeros(n), dtype=np.uint64)
    if not (len(c_start) == len(c_count) == len(c_lead) == n):
        raise AudioDecodeError("read_fragments: paths, starts, counts and leads must have the same length")
    if n and int((c_lead + c_count).max()) > int(want):
        raise AudioDecodeError("read_fragments: lead + count exceeds the clip length")
    failed = C.c_int64(-1)
    rc = lib.vmio_flac_read_fragments(c_paths, c_start.ctypes.data, c_count.ctypes.data, c_lead.ctypes.data, n,
                                      out.ctypes.data, int(want), max(1, int(workers)), C.byref(failed))
    if rc != 0:
        where = paths[failed.value] if 0 <= failed.value < n else "<batch>"
        _fail(rc, f"read_fragments({where})")
    return out


def read_flac(path, dtype="float64"):
    try:
        data = np.fromfile(path, dtype=np.uint8)
    except OSError as exc:
        raise AudioDecodeError(f"cannot read {path}: {exc}") from exc
    try:
        return decode_flac(data, dtype)
    except AudioDecodeError as exc:
        raise AudioDecodeError(f"{path}: {exc}") from None


def read_wav(path):
    """Integer-PCM WAV through the standard library; same scaling as soundfile (8-bit is unsigned in the file)."""
    import wave
    with wave.open(os.fspath(path), "rb") as handle:
        channels, width, rate, frames = handle.getnchannels(), handle.getsampwidth(), handle.getframerate(), handle.getnframes()
        raw = handle.readframes(frames)
    if width == 1:
        pcm = np.frombuffer(raw, dtype=np.uint8).astype(np.int32) - 128
    elif width == 2:
        pcm = np.frombuffer(raw, dtype="<i2").astype(np.int32)
    elif width == 3:
        b = np.frombuffer(raw, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
        pcm = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        pcm = np.where(pcm & 0x800000, pcm - (1 << 24), pcm)
    elif width == 4:
        pcm = np.frombuffer(raw, dtype="<i4")
    else:
        raise AudioDecodeError(f"{path}: unsupported WAV sample width {width}")
    samples = pcm.astype(np.float64) / float(1 << (8 * width - 1))
    if channels > 1:
        samples = samples.reshape(-1, channels)
    return samples, rate


def read(path):
    """``(samples, samplerate)`` of an audio file, like ``soundfile.read(path)``."""
    ext = os.path.splitext(os.fspath(path))[1].lower()
    if ext == ".flac":
        return read_flac(path)
    if ext == ".wav":
        return read_wav(path)
    try:
        import soundfile
    except ImportError as exc:
        raise AudioDecodeError(f"{path}: only .flac and .wav are decoded natively; other formats need the "
                               "`soundfile` package") from exc
    return soundfile.read(path)


def read_many(paths, reader=read, workers=None):
    """``[reader(p) for p in paths]`` on a thread pool (results in order).  ``paths`` may hold argument tuples, which
    are unpacked (``reader(*p)``).  The FLAC decoder runs outside the GIL, so threads scale with cores."""
    jobs = [p if isinstance(p, tuple) else (p,) for p in paths]
    if workers is None:
        workers = DEFAULT_WORKERS if DEFAULT_WORKERS is not None else min(len(jobs), os.cpu_count() or 1, 16)
    if workers <= 1 or len(jobs) <= 1:
        return [reader(*job) for job in jobs]
    with ThreadPoolExecutor(max_workers=workers) as pool:
        return list(pool.map(lambda job: reader(*job), jobs))
