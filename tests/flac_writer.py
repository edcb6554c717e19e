"""Test-side FLAC *encoder* (pure Python, slow, small inputs): produces streams that exercise every branch of the
decoder in voicemap_b200/csrc/vm_flac.c -- all subframe types, both Rice methods, escaped partitions, wasted bits, the
three stereo decorrelations, every block-size / sample-rate header form, ID3 tags around the stream.

Test infrastructure only (nothing in the product imports it).  It shares no code with the decoder; the two RFC 9639
appendix streams in tests/test_audio_io.py anchor the decoder independently of this writer.
"""
import hashlib

import numpy as np


class BitWriter:
    def __init__(self):
        self.acc = 0
        self.nbits = 0

    def put(self, value, n):
        if n:
            self.acc = (self.acc << n) | (int(value) & ((1 << n) - 1))
            self.nbits += n

    def put_unary(self, zeros):
        self.put(1, zeros + 1)

    def align(self):
        self.put(0, (-self.nbits) % 8)

    def bytes(self):
        assert self.nbits % 8 == 0
        return self.acc.to_bytes(self.nbits // 8, "big") if self.nbits else b""


def crc8(data):
    c = 0
    for byte in data:
        c ^= byte
        for _ in range(8):
            c = ((c << 1) ^ 0x07) & 0xff if c & 0x80 else (c << 1) & 0xff
    return c


def crc16(data):
    c = 0
    for byte in data:
        c ^= byte << 8
        for _ in range(8):
            c = ((c << 1) ^ 0x8005) & 0xffff if c & 0x8000 else (c << 1) & 0xffff
    return c


def _utf8_number(value):
    """FLAC's frame/sample number: UTF-8 extended to 36 bits (up to 7 bytes)."""
    if value < 0x80:
        return bytes([value])
    n = next(k for k in range(2, 8) if value < (1 << (5 * k + 1)) or k == 7)
    tail = [0x80 | ((value >> (6 * i)) & 0x3f) for i in range(n - 1)][::-1]
    lead = ((0xff << (8 - n)) & 0xff) | (value >> (6 * (n - 1)))
    return bytes([lead] + tail)


_BLOCK_CODES = {192: 1, 576: 2, 1152: 3, 2304: 4, 4608: 5, 256: 8, 512: 9, 1024: 10, 2048: 11, 4096: 12, 8192: 13,
                16384: 14, 32768: 15}
_RATE_CODES = {88200: 1, 176400: 2, 192000: 3, 8000: 4, 16000: 5, 22050: 6, 24000: 7, 32000: 8, 44100: 9, 48000: 10,
               96000: 11}
_BITS_CODES = {8: 1, 12: 2, 16: 4, 20: 5, 24: 6, 32: 7}


def _zigzag(v):
    return (v << 1) if v >= 0 else ((-v) << 1) - 1


def _rice_bits(values, k):
    return sum((_zigzag(v) >> k) + 1 + k for v in values)


def _write_residual(w, residual, blocksize, order, method, partition_order, escape_partitions):
    param_bits, escape = (4, 15) if method == 0 else (5, 31)
    w.put(method, 2)
    w.put(partition_order, 4)
    per = blocksize >> partition_order
    assert per << partition_order == blocksize and per >= order
    pos = 0
    for part in range(1 << partition_order):
        count = per - (order if part == 0 else 0)
        values = residual[pos:pos + count]
        pos += count
        if part in escape_partitions:
            raw = max([1] + [(v if v >= 0 else ~v).bit_length() + 1 for v in values]) if any(values) else 0
            w.put(escape, param_bits)
            w.put(raw, 5)
            for v in values:
                w.put(v, raw)
        else:
            k = min(range(escape), key=lambda kk: _rice_bits(values, kk)) if values else 0
            w.put(k, param_bits)
            for v in values:
                u = _zigzag(v)
                w.put_unary(u >> k)
                w.put(u, k)
    assert pos == len(residual)


def _lpc_coefficients(signal, order, precision):
    """Least-squares predictor quantised to `precision` bits with a non-negative shift."""
    x = np.asarray(signal, dtype=np.float64)
    if len(x) <= 2 * order or not np.any(x):
        return [0] * order, 0
    rows = np.stack([x[order - 1 - j:len(x) - 1 - j] for j in range(order)], axis=1)
    coef = np.linalg.lstsq(rows, x[order:], rcond=None)[0]
    top = max(np.abs(coef).max(), 1e-9)
    shift = int(np.clip(precision - 1 - np.ceil(np.log2(top)) - 1, 0, 15))
    q = np.clip(np.round(coef * (1 << shift)), -(1 << (precision - 1)), (1 << (precision - 1)) - 1).astype(int)
    return [int(v) for v in q], shift


def _write_subframe(w, signal, bps, kind, *, fixed_order=2, lpc_order=8, lpc_precision=12, method=0, partition_order=0,
                    escape_partitions=(), allow_wasted=True):
    signal = [int(v) for v in signal]
    n = len(signal)
    wasted = 0
    if allow_wasted and any(signal):
        while all((v >> wasted) & 1 == 0 for v in signal):
            wasted += 1
    if wasted:
        signal = [v >> wasted for v in signal]
        bps -= wasted
    if kind == "auto":
        kind = "constant" if len(set(signal)) == 1 else "fixed"
    if kind == "constant":
        assert len(set(signal)) == 1
        type_bits = 0
    elif kind == "verbatim":
        type_bits = 1
    elif kind == "fixed":
        type_bits = 8 + fixed_order
    elif kind == "lpc":
        type_bits = 31 + lpc_order
    else:
        raise ValueError(kind)
    w.put(0, 1)
    w.put(type_bits, 6)
    w.put(1 if wasted else 0, 1)
    if wasted:
        w.put_unary(wasted - 1)

    if kind == "constant":
        w.put(signal[0], bps)
    elif kind == "verbatim":
        for v in signal:
            w.put(v, bps)
    elif kind == "fixed":
        residual = list(signal)
        for _ in range(fixed_order):  # each pass shortens the list by one: n - order entries remain
            residual = [b - a for a, b in zip(residual, residual[1:])]
        for v in signal[:fixed_order]:
            w.put(v, bps)
        _write_residual(w, residual, n, fixed_order, method, partition_order, escape_partitions)
    else:
        coef, shift = _lpc_coefficients(signal, lpc_order, lpc_precision)
        for v in signal[:lpc_order]:
            w.put(v, bps)
        w.put(lpc_precision - 1, 4)
        w.put(shift, 5)
        for c in coef:
            w.put(c, lpc_precision)
        residual = [signal[i] - (sum(c * signal[i - 1 - j] for j, c in enumerate(coef)) >> shift)
                    for i in range(lpc_order, n)]
        _write_residual(w, residual, n, lpc_order, method, partition_order, escape_partitions)


def encode_frame(block, number, rate, bps, *, stereo="independent", variable=False, first_sample=0,
                 rate_from_streaminfo=False, bits_from_streaminfo=False, **subframe):
    """One frame from `block` (int array of shape (blocksize, channels))."""
    block = np.asarray(block).astype(object)
    n, channels = block.shape
    w = BitWriter()
    w.put(0x7ffc, 15)
    w.put(1 if variable else 0, 1)
    bs_code = _BLOCK_CODES.get(n, 6 if n <= 256 else 7)
    if rate_from_streaminfo:
        sr_code = 0
    elif rate in _RATE_CODES:
        sr_code = _RATE_CODES[rate]
    elif rate % 1000 == 0 and rate // 1000 < 256:
        sr_code = 12
    elif rate < 65536:
        sr_code = 13
    else:
        assert rate % 10 == 0
        sr_code = 14
    w.put(bs_code, 4)
    w.put(sr_code, 4)
    ch_code = {"independent": channels - 1, "left_side": 8, "side_right": 9, "mid_side": 10}[stereo]
    w.put(ch_code, 4)
    w.put(0 if bits_from_streaminfo else _BITS_CODES[bps], 3)
    w.put(0, 1)
    for byte in _utf8_number(first_sample if variable else number):
        w.put(byte, 8)
    if bs_code == 6:
        w.put(n - 1, 8)
    elif bs_code == 7:
        w.put(n - 1, 16)
    if sr_code == 12:
        w.put(rate // 1000, 8)
    elif sr_code == 13:
        w.put(rate, 16)
    elif sr_code == 14:
        w.put(rate // 10, 16)
    w.put(crc8(w.bytes()), 8)

    if stereo == "independent":
        signals = [(block[:, c], bps) for c in range(channels)]
    else:
        left, right = block[:, 0], block[:, 1]
        side = left - right
        if stereo == "left_side":
            signals = [(left, bps), (side, bps + 1)]
        elif stereo == "side_right":
            signals = [(side, bps + 1), (right, bps)]
        else:
            signals = [((left + right) >> 1, bps), (side, bps + 1)]
    for signal, width in signals:
        _write_subframe(w, signal, width, **subframe)
    w.align()
    body = w.bytes()
    return body + crc16(body).to_bytes(2, "big")


def encode_flac(samples, rate=16000, bps=16, blocksize=4096, *, kind="fixed", blocksizes=None, record_length=True,
                record_md5=True, extra_metadata=(), id3v2=b"", id3v1=False, **frame_options):
    """A complete FLAC stream.  `samples`: int array (frames,) or (frames, channels); `blocksizes`: explicit list
    of block sizes (variable-blocksize stream) instead of a fixed `blocksize`."""
    pcm = np.asarray(samples)
    if pcm.ndim == 1:
        pcm = pcm[:, None]
    total, channels = pcm.shape
    if blocksizes is None:
        sizes = [blocksize] * (total // blocksize) + ([total % blocksize] if total % blocksize else [])
        variable = False
    else:
        sizes = list(blocksizes)
        assert sum(sizes) == total
        variable = True
    frames = []
    start = 0
    for number, n in enumerate(sizes):
        frames.append(encode_frame(pcm[start:start + n], number, rate, bps, variable=variable, first_sample=start,
                                   kind=kind, **frame_options))
        start += n

    width = (bps + 7) // 8
    md5 = hashlib.md5(b"".join(int(v).to_bytes(width, "little", signed=True) for v in pcm.reshape(-1))).digest()
    info = BitWriter()
    info.put(min(sizes) if variable else blocksize, 16)
    info.put(max(sizes) if variable else blocksize, 16)
    info.put(min(map(len, frames), default=0), 24)
    info.put(max(map(len, frames), default=0), 24)
    info.put(rate, 20)
    info.put(channels - 1, 3)
    info.put(bps - 1, 5)
    info.put(total if record_length else 0, 36)
    blocks = [(0, info.bytes() + (md5 if record_md5 else bytes(16)))] + list(extra_metadata)
    out = bytearray(id3v2)
    out += b"fLaC"
    for i, (block_type, payload) in enumerate(blocks):
        out.append((0x80 if i == len(blocks) - 1 else 0) | block_type)
        out += len(payload).to_bytes(3, "big")
        out += payload
    for frame in frames:
        out += frame
    if id3v1:
        out += b"TAG" + bytes(125)
    return bytes(out)


def id3v2_tag(payload_size):
    size = bytes([(payload_size >> 21) & 0x7f, (payload_size >> 14) & 0x7f, (payload_size >> 7) & 0x7f,
                  payload_size & 0x7f])
    return b"ID3\x04\x00\x00" + size + bytes(payload_size)


# ------------------------------------------------------------------------------------------------ fast paths
_CRC16_TABLE = []
for _i in range(256):
    _c = _i << 8
    for _ in range(8):
        _c = ((_c << 1) ^ 0x8005) & 0xffff if _c & 0x8000 else (_c << 1) & 0xffff
    _CRC16_TABLE.append(_c)


def _crc16_fast(data):
    c = 0
    table = _CRC16_TABLE
    for byte in data:
        c = ((c << 8) & 0xffff) ^ table[(c >> 8) ^ byte]
    return c


def encode_flac_quick(pcm, rate=16000, blocksize=4096, constant=None):
    """16-bit mono stream for corpus-sized test data, built with array operations instead of per-sample bit writes:
    VERBATIM subframes (byte aligned for 16-bit samples) or, with `constant` = (value, total_samples), CONSTANT
    subframes (a whole file in a few hundred bytes)."""
    if constant is not None:
        value, total = constant
        pcm = None
    else:
        pcm = np.asarray(pcm).astype(np.int64)
        total = len(pcm)
    frames = []
    start = 0
    number = 0
    while start < total:
        n = min(blocksize, total - start)
        w = BitWriter()
        w.put(0x7ffc, 15)
        w.put(0, 1)
        bs_code = _BLOCK_CODES.get(n, 6 if n <= 256 else 7)
        w.put(bs_code, 4)
        w.put(_RATE_CODES[rate], 4)
        w.put(0, 4)
        w.put(_BITS_CODES[16], 3)
        w.put(0, 1)
        for byte in _utf8_number(number):
            w.put(byte, 8)
        if bs_code == 6:
            w.put(n - 1, 8)
        elif bs_code == 7:
            w.put(n - 1, 16)
        w.put(crc8(w.bytes()), 8)
        if pcm is None:
            body = w.bytes() + bytes([0x00]) + int(value).to_bytes(2, "big", signed=True)
        else:
            body = w.bytes() + bytes([0x02]) + pcm[start:start + n].astype(">i2").tobytes()
        frames.append(body + _crc16_fast(body).to_bytes(2, "big"))
        start += n
        number += 1
    if pcm is None:
        md5 = bytes(16)
    else:
        md5 = hashlib.md5(pcm.astype("<i2").tobytes()).digest()
    info = BitWriter()
    info.put(blocksize, 16)
    info.put(blocksize, 16)
    info.put(min(map(len, frames)), 24)
    info.put(max(map(len, frames)), 24)
    info.put(rate, 20)
    info.put(0, 3)
    info.put(15, 5)
    info.put(total, 36)
    return b"fLaC" + bytes([0x80]) + (34).to_bytes(3, "big") + info.bytes() + md5 + b"".join(frames)
