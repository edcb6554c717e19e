/* FLAC decoder behind include/voicemap_io.h (host side of the LibriSpeech batcher; plain C99, no dependencies).
 *
 * Replaces `soundfile.read` (libsndfile/libFLAC) at voicemap/librispeech.py:104,267.  Written from the format
 * description (RFC 9639): stream marker, metadata blocks, frames = header (CRC-8) + one subframe per channel +
 * padding + CRC-16; subframes are CONSTANT, VERBATIM, FIXED (order 0-4) or LPC (order 1-32) predictors with a
 * partitioned Rice-coded residual.  Integer exact: a FLAC stream has exactly one decoding.
 */
#include "../../include/voicemap_io.h"

#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define VMIO_VERSION 100
#define MAX_CHANNELS 8
#define MAX_BLOCK 65536

/* ------------------------------------------------------------------------------------------- checksums */
static uint8_t crc8_table[256];
static uint16_t crc16_table[8][256]; /* [k][b]: CRC of byte b followed by k zero bytes (slicing by 8) */

__attribute__((constructor)) static void build_crc_tables(void) {
    for (int i = 0; i < 256; ++i) {
        uint8_t c8 = (uint8_t)i;
        uint16_t c16 = (uint16_t)(i << 8);
        for (int b = 0; b < 8; ++b) {
            c8 = (uint8_t)((c8 & 0x80) ? ((c8 << 1) ^ 0x07) : (c8 << 1));         /* x^8 + x^2 + x + 1 */
            c16 = (uint16_t)((c16 & 0x8000) ? ((c16 << 1) ^ 0x8005) : (c16 << 1)); /* x^16 + x^15 + x^2 + 1 */
        }
        crc8_table[i] = c8;
        crc16_table[0][i] = c16;
    }
    for (int k = 1; k < 8; ++k)
        for (int i = 0; i < 256; ++i) {
            const uint16_t c = crc16_table[k - 1][i];
            crc16_table[k][i] = (uint16_t)((c << 8) ^ crc16_table[0][c >> 8]);
        }
}

static uint8_t crc8(const uint8_t* p, size_t n) {
    uint8_t c = 0;
    while (n--) c = crc8_table[c ^ *p++];
    return c;
}

static uint16_t crc16(const uint8_t* p, size_t n) {
    uint16_t c = 0;
    for (; n >= 8; n -= 8, p += 8) /* the 16-bit state only meets the first two bytes; the other six are independent lookups */
        c = (uint16_t)(crc16_table[7][p[0] ^ (c >> 8)] ^ crc16_table[6][p[1] ^ (c & 0xff)] ^ crc16_table[5][p[2]] ^
                       crc16_table[4][p[3]] ^ crc16_table[3][p[4]] ^ crc16_table[2][p[5]] ^ crc16_table[1][p[6]] ^
                       crc16_table[0][p[7]]);
    while (n--) c = (uint16_t)((c << 8) ^ crc16_table[0][(c >> 8) ^ *p++]);
    return c;
}

/* ------------------------------------------------------------------------------------------- bit reader
 * MSB first.  `acc` is left aligned: its top `nbits` bits are the next bits of the stream (bits below them are either
 * zero or a correct preview of what follows -- the 8-byte refill ORs in a whole word and only counts whole bytes).
 * Reading past the end sets `eof` and yields zeros; callers check `eof` once per subframe / frame. */
typedef struct {
    const uint8_t* p;
    size_t len, pos;
    uint64_t acc;
    int nbits;
    int eof;
} bits_t;

static inline void refill(bits_t* b) { /* afterwards nbits >= 56 unless the data ends */
    if (b->pos + 8 <= b->len) {
        const uint8_t* q = b->p + b->pos;
        const uint64_t w = ((uint64_t)q[0] << 56) | ((uint64_t)q[1] << 48) | ((uint64_t)q[2] << 40) | ((uint64_t)q[3] << 32) |
                           ((uint64_t)q[4] << 24) | ((uint64_t)q[5] << 16) | ((uint64_t)q[6] << 8) | (uint64_t)q[7];
        b->acc |= w >> b->nbits;           /* nbits <= 63 here */
        b->pos += (size_t)((63 - b->nbits) >> 3);
        b->nbits |= 56;
    } else {
        while (b->nbits <= 56 && b->pos < b->len) {
            b->acc |= (uint64_t)b->p[b->pos++] << (56 - b->nbits);
            b->nbits += 8;
        }
    }
}

static inline uint32_t read_bits(bits_t* b, int n) { /* 0 <= n <= 32 */
    if (n == 0) return 0;
    if (b->nbits < n) {
        refill(b);
        if (b->nbits < n) { /* end of data: zeros */
            b->eof = 1;
            b->acc = 0;
            b->nbits = 0;
            return 0;
        }
    }
    const uint32_t v = (uint32_t)(b->acc >> (64 - n));
    b->acc <<= n;
    b->nbits -= n;
    return v;
}

static inline int64_t read_signed(bits_t* b, int n) { /* 1 <= n <= 33 */
    uint64_t v = 0;
    if (n > 32) {
        v = read_bits(b, n - 32);
        v = (v << 32) | read_bits(b, 32);
    } else {
        v = read_bits(b, n);
    }
    const uint64_t sign = 1ull << (n - 1);
    return (int64_t)((v ^ sign) - sign);
}

static inline int64_t read_unary(bits_t* b) { /* number of 0 bits before the next 1 bit; -1 at end of data */
    int64_t count = 0;
    for (;;) {
        if (b->acc != 0) {
            const int lz = __builtin_clzll(b->acc);
            if (lz < b->nbits) {
                b->acc <<= lz;
                b->acc <<= 1; /* two shifts: lz + 1 may be 64 */
                b->nbits -= lz + 1;
                return count + lz;
            }
        }
        /* all counted bits are zero: take them and fetch more */
        count += b->nbits;
        b->acc = 0;
        b->nbits = 0;
        refill(b);
        if (b->nbits == 0) {
            b->eof = 1;
            return -1;
        }
    }
}

/* first byte not yet touched by the reader, once the reader is byte aligned (nbits is a multiple of 8) */
static inline size_t byte_position(const bits_t* b) { return b->pos - (size_t)(b->nbits >> 3); }

/* ------------------------------------------------------------------------------------------- metadata */
static uint32_t be(const uint8_t* p, int n) {
    uint32_t v = 0;
    for (int i = 0; i < n; ++i) v = (v << 8) | p[i];
    return v;
}

/* Offset of the first frame, or a negative error.  `complete` = the whole stream is in `data` (a probe of the first
 * bytes of a file may stop inside a later metadata block once STREAMINFO has been seen). */
static int64_t parse_metadata(const uint8_t* data, size_t len, vmio_flac_info* info, int complete) {
    size_t off = 0;
    if (len >= 10 && memcmp(data, "ID3", 3) == 0) { /* ID3v2 tag in front of the stream */
        const size_t tag = ((size_t)(data[6] & 0x7f) << 21) | ((size_t)(data[7] & 0x7f) << 14) |
                           ((size_t)(data[8] & 0x7f) << 7) | (size_t)(data[9] & 0x7f);
        off = 10 + tag;
    }
    if (len < off + 4) return complete ? VMIO_ERR_NOT_FLAC : VMIO_ERR_TRUNCATED;
    if (memcmp(data + off, "fLaC", 4) != 0) return VMIO_ERR_NOT_FLAC;
    off += 4;
    int seen_streaminfo = 0;
    for (;;) {
        if (len < off + 4) return seen_streaminfo && !complete ? (int64_t)off : VMIO_ERR_TRUNCATED;
        const int last = data[off] >> 7;
        const int type = data[off] & 0x7f;
        const size_t size = be(data + off + 1, 3);
        off += 4;
        if (!seen_streaminfo) {
            if (type != 0 || size < 34) return VMIO_ERR_NOT_FLAC; /* STREAMINFO must come first */
            if (len < off + 34) return VMIO_ERR_TRUNCATED;
            const uint8_t* s = data + off;
            info->min_blocksize = be(s, 2);
            info->max_blocksize = be(s + 2, 2);
            info->sample_rate = be(s + 10, 3) >> 4;
            info->channels = ((s[12] >> 1) & 7) + 1;
            info->bits_per_sample = (((s[12] & 1) << 4) | (s[13] >> 4)) + 1;
            info->total_samples = ((uint64_t)(s[13] & 0x0f) << 32) | be(s + 14, 4);
            memcpy(info->md5, s + 18, 16);
            seen_streaminfo = 1;
        }
        if (len < off + size) return !complete ? (int64_t)len : VMIO_ERR_TRUNCATED;
        off += size;
        if (last) return (int64_t)off;
    }
}

/* ------------------------------------------------------------------------------------------- subframes */
/* sum_j coef[j] * p[-1-j]: the prediction for the sample at p.  The product that needs the newest sample is added last,
 * so the serial chain from one sample to the next is a multiply, an add, a shift and an add. */
static inline int64_t lpc_dot(const int64_t* p, const int64_t* coef, int order) {
    int64_t acc = 0;
    for (int j = order - 1; j >= 1; --j) acc += coef[j] * p[-1 - j];
    return acc + coef[0] * p[-1];
}

/* One partition of `count` Rice-coded residuals with parameter k, written to out[i...]; PREDICTION is an expression in
 * `p` (the address the sample goes to) that is added to the residual before it is stored.  A macro and not a function
 * because the variants below differ in a compile-time constant (the predictor order) inside the innermost loop. */
#define RICE_PARTITION(PREDICTION)                                                                      \
    for (uint32_t j = 0; j < count; ++j, ++i) {                                                         \
        uint64_t u;                                                                                     \
        if (b->nbits < 48) refill(b);                                                                   \
        const int lz = b->acc ? __builtin_clzll(b->acc) : 64;                                           \
        if (lz + 1 + (int)k <= b->nbits) { /* quotient, stop bit and remainder are all in the window */ \
            uint64_t w = (b->acc << lz) << 1;                                                           \
            u = ((uint64_t)lz << k) | ((w >> 1) >> (63 - k));                                           \
            b->acc = w << k;                                                                            \
            b->nbits -= lz + 1 + (int)k;                                                                \
        } else {                                                                                        \
            const int64_t q = read_unary(b);                                                            \
            if (q < 0) return VMIO_ERR_TRUNCATED;                                                       \
            u = ((uint64_t)q << k) | read_bits(b, (int)k);                                              \
        }                                                                                               \
        int64_t* const p = out + i;                                                                     \
        *p = ((int64_t)(u >> 1) ^ -(int64_t)(u & 1)) /* zig-zag: 0,-1,1,-2,... */ + (PREDICTION);       \
    }
#define FUSED_CASE(ORDER) case ORDER: RICE_PARTITION(lpc_dot(p, coef, ORDER) >> shift) break;

/* Residuals of one subframe into out[order ... blocksize).  With `coef` != NULL (LPC orders 1-12, what encoders emit for
 * speech) every sample is completed on the spot -- residual + (prediction >> shift) -- in the same loop that decodes its
 * residual: the bit reader's dependency chain and the predictor's are independent, so the core overlaps them (measured
 * 14 % faster on order-8 streams than decoding all residuals first and running the predictor over them afterwards).
 * With `coef` == NULL the plain residuals are stored and the caller applies its predictor. */
static int decode_residual(bits_t* b, int64_t* out, uint32_t blocksize, int order, const int64_t* coef, int shift) {
    const int method = (int)read_bits(b, 2);
    if (method > 1) return VMIO_ERR_HEADER;
    const int param_bits = method == 0 ? 4 : 5;
    const uint32_t escape = method == 0 ? 15u : 31u;
    const int porder = (int)read_bits(b, 4);
    const uint32_t partitions = 1u << porder;
    if ((blocksize & (partitions - 1)) != 0 || (blocksize >> porder) < (uint32_t)order) return VMIO_ERR_RESIDUAL;
    uint32_t i = (uint32_t)order;
    for (uint32_t part = 0; part < partitions; ++part) {
        const uint32_t count = (blocksize >> porder) - (part == 0 ? (uint32_t)order : 0u);
        const uint32_t k = read_bits(b, param_bits);
        if (k == escape) { /* residuals stored as plain `raw`-bit numbers */
            const int raw = (int)read_bits(b, 5);
            for (uint32_t j = 0; j < count; ++j, ++i) {
                const int64_t r = raw ? read_signed(b, raw) : 0;
                out[i] = coef ? r + (lpc_dot(out + i, coef, order) >> shift) : r;
            }
        } else {
            switch (coef ? order : 0) {
                case 0: RICE_PARTITION(0) break;
                FUSED_CASE(1) FUSED_CASE(2) FUSED_CASE(3) FUSED_CASE(4) FUSED_CASE(5) FUSED_CASE(6)
                FUSED_CASE(7) FUSED_CASE(8) FUSED_CASE(9) FUSED_CASE(10) FUSED_CASE(11) FUSED_CASE(12)
                default: return VMIO_ERR_ARG; /* callers fuse orders 1-12 only */
            }
        }
        if (b->eof) return VMIO_ERR_TRUNCATED;
    }
    return VMIO_OK;
}
#undef FUSED_CASE
#undef RICE_PARTITION
#define FUSED_LPC_MAX_ORDER 12

static int decode_subframe(bits_t* b, int64_t* s, uint32_t n, int bps) {
    if (read_bits(b, 1) != 0) return VMIO_ERR_HEADER;
    const int type = (int)read_bits(b, 6);
    int wasted = 0;
    if (read_bits(b, 1)) {
        const int64_t z = read_unary(b);
        if (z < 0) return VMIO_ERR_TRUNCATED;
        wasted = (int)z + 1;
    }
    bps -= wasted;
    if (bps <= 0) return VMIO_ERR_HEADER;

    if (type == 0) { /* CONSTANT */
        const int64_t v = read_signed(b, bps);
        for (uint32_t i = 0; i < n; ++i) s[i] = v;
    } else if (type == 1) { /* VERBATIM */
        for (uint32_t i = 0; i < n; ++i) s[i] = read_signed(b, bps);
    } else if (type >= 8 && type <= 12) { /* FIXED predictor of order type-8 */
        const int order = type - 8;
        if ((uint32_t)order > n) return VMIO_ERR_HEADER;
        for (int i = 0; i < order; ++i) s[i] = read_signed(b, bps);
        const int rc = decode_residual(b, s, n, order, NULL, 0);
        if (rc) return rc;
        switch (order) {
            case 0: break;
            case 1: for (uint32_t i = 1; i < n; ++i) s[i] += s[i - 1]; break;
            case 2: for (uint32_t i = 2; i < n; ++i) s[i] += 2 * s[i - 1] - s[i - 2]; break;
            case 3: for (uint32_t i = 3; i < n; ++i) s[i] += 3 * s[i - 1] - 3 * s[i - 2] + s[i - 3]; break;
            case 4:
                for (uint32_t i = 4; i < n; ++i) s[i] += 4 * s[i - 1] - 6 * s[i - 2] + 4 * s[i - 3] - s[i - 4];
                break;
        }
    } else if (type >= 32) { /* LPC of order type-31 */
        const int order = type - 31;
        if ((uint32_t)order > n) return VMIO_ERR_HEADER;
        for (int i = 0; i < order; ++i) s[i] = read_signed(b, bps);
        const int precision = (int)read_bits(b, 4) + 1;
        if (precision == 16) return VMIO_ERR_HEADER;
        const int shift = (int)read_signed(b, 5);
        if (shift < 0) return VMIO_ERR_HEADER;
        int64_t coef[32];
        for (int j = 0; j < order; ++j) coef[j] = read_signed(b, precision);
        const int fused = order <= FUSED_LPC_MAX_ORDER;
        const int rc = decode_residual(b, s, n, order, fused ? coef : NULL, shift);
        if (rc) return rc;
        if (!fused) /* long predictors: residuals first, then the recurrence */
            for (uint32_t i = (uint32_t)order; i < n; ++i) s[i] += lpc_dot(s + i, coef, order) >> shift;
    } else {
        return VMIO_ERR_HEADER; /* reserved subframe type */
    }
    if (wasted)
        for (uint32_t i = 0; i < n; ++i) s[i] *= (int64_t)1 << wasted;
    return b->eof ? VMIO_ERR_TRUNCATED : VMIO_OK;
}

/* ------------------------------------------------------------------------------------------- frames */
static const uint32_t kRates[12] = {0, 88200, 176400, 192000, 8000, 16000, 22050, 24000, 32000, 44100, 48000, 96000};
static const int kBits[8] = {0, 8, 12, -1, 16, 20, 24, 32};

/* Scratch samples per channel: the stream's declared maximum block size (a few thousand samples, so the buffer comes
 * from the allocator's arena instead of a fresh mmap per call, which serialises threads on the process's mmap lock). */
static size_t scratch_stride(const vmio_flac_info* info) {
    return info->max_blocksize >= 16 && info->max_blocksize <= MAX_BLOCK ? info->max_blocksize : MAX_BLOCK;
}

/* Decodes one frame starting at data[*offset]; channel c lands in chan[c * stride ...] with stride = scratch_stride().  `position_out` (optional)
 * receives the index of the frame's first sample as coded in its header. */
static int decode_frame(const uint8_t* data, size_t len, size_t* offset, const vmio_flac_info* info, int64_t* chan,
                        uint32_t* blocksize_out, uint64_t* position_out) {
    const size_t start = *offset;
    const size_t stride = scratch_stride(info);
    if (len - start < 6) return VMIO_ERR_TRUNCATED;
    bits_t b = {data, len, start, 0, 0, 0};
    if (read_bits(&b, 15) != 0x7ffc) return VMIO_ERR_SYNC; /* 14 sync bits + reserved 0 */
    const int variable = (int)read_bits(&b, 1);            /* blocking strategy: what the coded number counts */
    const int bs_code = (int)read_bits(&b, 4);
    const int sr_code = (int)read_bits(&b, 4);
    const int ch_code = (int)read_bits(&b, 4);
    const int bits_code = (int)read_bits(&b, 3);
    if (read_bits(&b, 1) != 0) return VMIO_ERR_HEADER;
    if (bs_code == 0 || sr_code == 15 || ch_code > 10 || kBits[bits_code] < 0) return VMIO_ERR_HEADER;

    /* frame / sample number, UTF-8 style: up to 7 bytes (36 bits) */
    const uint32_t lead = read_bits(&b, 8);
    uint64_t number = lead;
    int follow = 0;
    if (lead & 0x80) {
        uint32_t m = lead;
        while (m & 0x80) { ++follow; m <<= 1; }
        if (follow < 2 || follow > 7) return VMIO_ERR_HEADER;
        number = lead & (0x7fu >> follow);
        --follow;
        for (int i = 0; i < follow; ++i) {
            const uint32_t next = read_bits(&b, 8);
            if ((next & 0xc0) != 0x80) return VMIO_ERR_HEADER;
            number = (number << 6) | (next & 0x3f);
        }
    }

    uint32_t blocksize;
    if (bs_code == 1) blocksize = 192;
    else if (bs_code <= 5) blocksize = 576u << (bs_code - 2);
    else if (bs_code == 6) blocksize = read_bits(&b, 8) + 1;
    else if (bs_code == 7) blocksize = read_bits(&b, 16) + 1;
    else blocksize = 256u << (bs_code - 8);

    uint32_t rate = sr_code < 12 ? kRates[sr_code] : 0;
    if (sr_code == 12) rate = read_bits(&b, 8) * 1000u;
    else if (sr_code == 13) rate = read_bits(&b, 16);
    else if (sr_code == 14) rate = read_bits(&b, 16) * 10u;
    (void)rate; /* the stream's rate is STREAMINFO's; per-frame rates are not cross-checked */

    if (blocksize > stride) return VMIO_ERR_HEADER; /* larger than STREAMINFO's maximum block size */
    const size_t header_end = byte_position(&b);
    const uint32_t want_crc8 = read_bits(&b, 8);
    if (b.eof) return VMIO_ERR_TRUNCATED;
    if (crc8(data + start, header_end - start) != want_crc8) return VMIO_ERR_CRC8;

    const int bps = bits_code ? kBits[bits_code] : (int)info->bits_per_sample;
    const int channels = ch_code < 8 ? ch_code + 1 : 2;
    if ((uint32_t)channels != info->channels || (uint32_t)bps != info->bits_per_sample) return VMIO_ERR_HEADER;

    for (int c = 0; c < channels; ++c) {
        /* the side channel of a decorrelated pair carries one extra bit */
        const int side = (ch_code == 8 && c == 1) || (ch_code == 9 && c == 0) || (ch_code == 10 && c == 1);
        const int rc = decode_subframe(&b, chan + (size_t)c * stride, blocksize, bps + side);
        if (rc) return rc;
    }
    read_bits(&b, b.nbits & 7); /* zero padding to the byte boundary */
    const size_t body_end = byte_position(&b);
    const uint32_t want_crc16 = read_bits(&b, 16);
    if (b.eof) return VMIO_ERR_TRUNCATED;
    if (crc16(data + start, body_end - start) != want_crc16) return VMIO_ERR_CRC16;

    int64_t* c0 = chan;
    int64_t* c1 = chan + stride;
    if (ch_code == 8) { /* left, side = left - right */
        for (uint32_t i = 0; i < blocksize; ++i) c1[i] = c0[i] - c1[i];
    } else if (ch_code == 9) { /* side, right */
        for (uint32_t i = 0; i < blocksize; ++i) c0[i] += c1[i];
    } else if (ch_code == 10) { /* mid = (left + right) >> 1, side = left - right: side's parity restores the lost bit */
        for (uint32_t i = 0; i < blocksize; ++i) {
            const int64_t side = c1[i];
            const int64_t mid = c0[i] * 2 + (side & 1);
            c0[i] = (mid + side) >> 1;
            c1[i] = (mid - side) >> 1;
        }
    }
    *offset = body_end + 2;
    *blocksize_out = blocksize;
    if (position_out) *position_out = variable ? number : number * (uint64_t)info->max_blocksize;
    return VMIO_OK;
}

/* ------------------------------------------------------------------------------------------- entry points */
int vmio_version(void) { return VMIO_VERSION; }

const char* vmio_error_string(int code) {
    switch (code) {
        case VMIO_OK: return "ok";
        case VMIO_ERR_ARG: return "bad argument";
        case VMIO_ERR_NOT_FLAC: return "not a FLAC stream (no fLaC marker / STREAMINFO)";
        case VMIO_ERR_TRUNCATED: return "stream is truncated";
        case VMIO_ERR_SYNC: return "lost frame synchronisation";
        case VMIO_ERR_HEADER: return "invalid frame or subframe header";
        case VMIO_ERR_CRC8: return "frame header CRC-8 mismatch";
        case VMIO_ERR_CRC16: return "frame CRC-16 mismatch";
        case VMIO_ERR_UNSUPPORTED: return "unsupported stream (more than 8 channels or 32 bits per sample)";
        case VMIO_ERR_CAPACITY: return "output buffer too small";
        case VMIO_ERR_RESIDUAL: return "inconsistent residual partitioning";
        case VMIO_ERR_NOMEM: return "out of memory";
        case VMIO_ERR_IO: return "file could not be read";
        case VMIO_ERR_SHORT: return "file holds fewer samples than requested (stale index?)";
        default: return "unknown error";
    }
}

int vmio_flac_probe(const uint8_t* data, size_t len, vmio_flac_info* info) {
    if (!data || !info || len == 0) return VMIO_ERR_ARG;
    memset(info, 0, sizeof *info);
    const int64_t off = parse_metadata(data, len, info, 0);
    return off < 0 ? (int)off : VMIO_OK;
}

int64_t vmio_flac_decode(const uint8_t* data, size_t len, int32_t* out_i32, double* out_f64,
                         uint64_t capacity_frames, vmio_flac_info* info_out) {
    if (!data || len == 0) return VMIO_ERR_ARG;
    vmio_flac_info info;
    memset(&info, 0, sizeof info);
    const int64_t first = parse_metadata(data, len, &info, 1);
    if (info_out) *info_out = info;
    if (first < 0) return first;
    if (info.channels > MAX_CHANNELS || info.bits_per_sample > 32 || info.bits_per_sample < 4) return VMIO_ERR_UNSUPPORTED;

    int64_t* chan = (int64_t*)malloc(sizeof(int64_t) * scratch_stride(&info) * info.channels);
    if (!chan) return VMIO_ERR_NOMEM;
    const int store = out_i32 != NULL || out_f64 != NULL;
    const double scale = 1.0 / (double)(1ull << (info.bits_per_sample - 1));
    const uint32_t nch = info.channels;
    size_t off = (size_t)first;
    uint64_t done = 0;
    int64_t rc = VMIO_OK;
    while (off < len) {
        if (info.total_samples && done >= info.total_samples) break;           /* anything after that is not audio */
        if (len - off >= 3 && memcmp(data + off, "TAG", 3) == 0) break;         /* ID3v1 tag at the end */
        uint32_t n = 0;
        rc = decode_frame(data, len, &off, &info, chan, &n, NULL);
        if (rc) break;
        if (store) {
            if (done + n > capacity_frames) { rc = VMIO_ERR_CAPACITY; break; }
            if (nch == 1) {
                if (out_i32) for (uint32_t i = 0; i < n; ++i) out_i32[done + i] = (int32_t)chan[i];
                if (out_f64) for (uint32_t i = 0; i < n; ++i) out_f64[done + i] = (double)chan[i] * scale;
            } else for (uint32_t c = 0; c < nch; ++c) {
                const int64_t* src = chan + (size_t)c * scratch_stride(&info);
                if (out_i32) {
                    int32_t* dst = out_i32 + done * nch + c;
                    for (uint32_t i = 0; i < n; ++i) dst[(size_t)i * nch] = (int32_t)src[i];
                }
                if (out_f64) {
                    double* dst = out_f64 + done * nch + c;
                    for (uint32_t i = 0; i < n; ++i) dst[(size_t)i * nch] = (double)src[i] * scale;
                }
            }
        }
        done += n;
    }
    free(chan);
    return rc ? rc : (int64_t)done;
}

/* First offset >= from where a frame decodes cleanly (sync code, header CRC-8, body, CRC-16), or `len`. */
static size_t find_frame(const uint8_t* data, size_t len, size_t from, const vmio_flac_info* info, int64_t* chan,
                         uint64_t* position) {
    for (size_t at = from; at + 6 <= len; ++at) {
        if (data[at] != 0xff || (data[at + 1] & 0xfe) != 0xf8) continue;
        size_t end = at;
        uint32_t n = 0;
        if (decode_frame(data, len, &end, info, chan, &n, position) == VMIO_OK) return at;
    }
    return len;
}

int64_t vmio_flac_decode_range(const uint8_t* data, size_t len, uint64_t start, uint64_t count, int32_t* out_i32,
                               double* out_f64, vmio_flac_info* info_out) {
    if (!data || len == 0 || (!out_i32 && !out_f64 && count)) return VMIO_ERR_ARG;
    vmio_flac_info info;
    memset(&info, 0, sizeof info);
    const int64_t first = parse_metadata(data, len, &info, 1);
    if (info_out) *info_out = info;
    if (first < 0) return first;
    if (info.channels > MAX_CHANNELS || info.bits_per_sample > 32 || info.bits_per_sample < 4) return VMIO_ERR_UNSUPPORTED;
    if (count == 0) return 0;
    int64_t* chan = (int64_t*)malloc(sizeof(int64_t) * scratch_stride(&info) * info.channels);
    if (!chan) return VMIO_ERR_NOMEM;

    /* Frames carry no length, so the way to skip audio is to jump: guess a byte offset from the average compression
     * ratio, resynchronise on the next frame that decodes cleanly and read its position; if that is already past
     * `start`, guess further back.  Without a recorded total length the stream is walked from its first frame. */
    size_t off = (size_t)first;
    uint64_t position = 0;
    if (info.total_samples && info.max_blocksize && start > 0) {
        uint64_t back = info.max_blocksize;
        for (int attempt = 0; attempt < 6; ++attempt, back *= 4) {
            if (start <= back) break;
            const double fraction = (double)(start - back) / (double)info.total_samples;
            if (fraction >= 1.0) break;
            const size_t guess = (size_t)first + (size_t)(fraction * (double)(len - (size_t)first));
            uint64_t p = 0;
            const size_t at = find_frame(data, len, guess, &info, chan, &p);
            if (at < len && p <= start) {
                off = at;
                position = p;
                break;
            }
        }
    }

    const double scale = 1.0 / (double)(1ull << (info.bits_per_sample - 1));
    const uint32_t nch = info.channels;
    const uint64_t stop = start + count;
    uint64_t written = 0;
    int64_t rc = VMIO_OK;
    while (off < len && position < stop) {
        if (info.total_samples && position >= info.total_samples) break;
        if (len - off >= 3 && memcmp(data + off, "TAG", 3) == 0) break;
        uint32_t n = 0;
        rc = decode_frame(data, len, &off, &info, chan, &n, NULL);
        if (rc) break;
        const uint64_t lo = position > start ? position : start;
        const uint64_t hi = position + n < stop ? position + n : stop;
        for (uint64_t t = lo; t < hi; ++t)
            for (uint32_t c = 0; c < nch; ++c) {
                const int64_t v = chan[(size_t)c * scratch_stride(&info) + (size_t)(t - position)];
                if (out_i32) out_i32[(t - start) * nch + c] = (int32_t)v;
                if (out_f64) out_f64[(t - start) * nch + c] = (double)v * scale;
            }
        if (hi > lo) written = hi - start;
        position += n;
    }
    free(chan);
    return rc ? rc : (int64_t)written;
}

static int64_t slurp(const char* path, size_t limit, uint8_t** out) {
    FILE* f = fopen(path, "rb");
    if (!f) return VMIO_ERR_IO;
    size_t cap = limit ? limit : (1u << 20), n = 0;
    uint8_t* buf = (uint8_t*)malloc(cap);
    if (!buf) { fclose(f); return VMIO_ERR_NOMEM; }
    for (;;) {
        n += fread(buf + n, 1, cap - n, f);
        if (n < cap || limit) break;
        uint8_t* grown = (uint8_t*)realloc(buf, cap * 2);
        if (!grown) { free(buf); fclose(f); return VMIO_ERR_NOMEM; }
        buf = grown;
        cap *= 2;
    }
    const int bad = ferror(f);
    fclose(f);
    if (bad) { free(buf); return VMIO_ERR_IO; }
    *out = buf;
    return (int64_t)n;
}

int64_t vmio_flac_read_file(const char* path, int32_t* out_i32, double* out_f64, uint64_t capacity_frames,
                            vmio_flac_info* info) {
    if (!path) return VMIO_ERR_ARG;
    uint8_t* buf = NULL;
    const int64_t n = slurp(path, 0, &buf);
    if (n < 0) return n;
    const int64_t rc = n == 0 ? VMIO_ERR_NOT_FLAC : vmio_flac_decode(buf, (size_t)n, out_i32, out_f64, capacity_frames, info);
    free(buf);
    return rc;
}

int vmio_flac_probe_file(const char* path, vmio_flac_info* info) {
    if (!path || !info) return VMIO_ERR_ARG;
    uint8_t* buf = NULL;
    int64_t n = slurp(path, 4096, &buf);
    if (n < 0) return (int)n;
    int rc = n == 0 ? VMIO_ERR_NOT_FLAC : vmio_flac_probe(buf, (size_t)n, info);
    free(buf);
    if (rc == VMIO_ERR_TRUNCATED || rc == VMIO_ERR_NOT_FLAC) { /* a large ID3v2 tag can push STREAMINFO past 4 KB */
        n = slurp(path, 0, &buf);
        if (n < 0) return (int)n;
        rc = n == 0 ? VMIO_ERR_NOT_FLAC : vmio_flac_probe(buf, (size_t)n, info);
        free(buf);
    }
    return rc;
}

/* ------------------------------------------------------------------------------------------- batch of fragments */
typedef struct {
    const char* const* paths;
    const uint64_t* start;
    const uint64_t* count;
    const uint64_t* lead;
    size_t n, want;
    double* out;
    size_t next;         /* next row to claim (atomic) */
    int error;           /* first error, 0 while none (atomic) */
    int64_t failed;      /* row of the first error */
} batch_t;

static int read_fragment(const batch_t* b, size_t i) {
    double* row = b->out + i * b->want;
    const uint64_t lead = b->lead ? b->lead[i] : 0, count = b->count[i];
    if (lead + count > b->want) return VMIO_ERR_ARG;
    memset(row, 0, sizeof(double) * lead);
    memset(row + lead + count, 0, sizeof(double) * (b->want - lead - count));
    if (count == 0) return VMIO_OK;
    uint8_t* buf = NULL;
    const int64_t size = slurp(b->paths[i], 0, &buf);
    if (size < 0) return (int)size;
    vmio_flac_info info;
    int64_t got = size == 0 ? VMIO_ERR_NOT_FLAC : vmio_flac_probe(buf, (size_t)size, &info);
    if (got == VMIO_OK) got = info.channels == 1 ? vmio_flac_decode_range(buf, (size_t)size, b->start[i], count, NULL, row + lead, NULL)
                                                 : VMIO_ERR_UNSUPPORTED;
    free(buf);
    if (got < 0) return (int)got;
    return (uint64_t)got == count ? VMIO_OK : VMIO_ERR_SHORT;
}

static void* batch_worker(void* arg) {
    batch_t* b = (batch_t*)arg;
    for (;;) {
        const size_t i = __atomic_fetch_add(&b->next, 1, __ATOMIC_RELAXED);
        if (i >= b->n || __atomic_load_n(&b->error, __ATOMIC_RELAXED)) break;
        const int rc = read_fragment(b, i);
        if (rc) {
            int none = 0;
            if (__atomic_compare_exchange_n(&b->error, &none, rc, 0, __ATOMIC_RELAXED, __ATOMIC_RELAXED))
                __atomic_store_n(&b->failed, (int64_t)i, __ATOMIC_RELAXED);
        }
    }
    return NULL;
}

int vmio_flac_read_fragments(const char* const* paths, const uint64_t* start, const uint64_t* count, const uint64_t* lead,
                             size_t n, double* out, size_t want, int threads, int64_t* failed_row) {
    if (failed_row) *failed_row = -1;
    if (n == 0) return VMIO_OK;
    if (!paths || !start || !count || !out) return VMIO_ERR_ARG;
    batch_t b = {paths, start, count, lead, n, want, out, 0, 0, -1};
    if (threads > 64) threads = 64;
    if ((size_t)threads > n) threads = (int)n;
    pthread_t ids[64];
    int started = 0;
    for (int t = 1; t < threads; ++t) /* the calling thread is worker 0 */
        if (pthread_create(&ids[started], NULL, batch_worker, &b) == 0) ++started;
    batch_worker(&b);
    for (int t = 0; t < started; ++t) pthread_join(ids[t], NULL);
    if (b.error && failed_row) *failed_row = b.failed;
    return b.error;
}

