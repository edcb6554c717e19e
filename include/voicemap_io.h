/* voicemap_io -- C ABI of the host-side audio decoder used by the LibriSpeech batcher.
 *
 * The reference decodes every LibriSpeech utterance with `soundfile.read(path)` (libsndfile -> libFLAC):
 * voicemap/librispeech.py:104 (items) and :267 (indexing).  Neither library is in this image, so the batcher
 * would have no way to read the corpus; this is a from-scratch FLAC decoder (the format of RFC 9639: fixed and LPC
 * predictors, partitioned Rice residuals, inter-channel decorrelation, wasted bits, frame CRCs) with the result
 * convention of `soundfile.read`: float64 samples = integer PCM / 2^(bits-1), interleaved by channel.
 *
 * Host only (plain C, no CUDA), thread safe (no global state), never allocates on behalf of the caller beyond
 * short-lived scratch that is freed before returning, never throws: 0 / a count on success, a negative VMIO_ERR_*
 * code on failure.  Python binds it with ctypes (voicemap_b200/audio_io.py), which releases the GIL around each
 * call so that a thread pool decodes a batch of files in parallel.
 */
#ifndef VOICEMAP_IO_H
#define VOICEMAP_IO_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VMIO_OK 0
#define VMIO_ERR_ARG (-1)         /* null pointer / zero length */
#define VMIO_ERR_NOT_FLAC (-2)    /* no "fLaC" marker (after an optional ID3v2 tag) or no STREAMINFO block */
#define VMIO_ERR_TRUNCATED (-3)   /* stream ends inside a metadata block or a frame */
#define VMIO_ERR_SYNC (-4)        /* frame sync code not found where a frame must start */
#define VMIO_ERR_HEADER (-5)      /* reserved / invalid field in a frame or subframe header */
#define VMIO_ERR_CRC8 (-6)        /* frame header checksum mismatch */
#define VMIO_ERR_CRC16 (-7)       /* frame checksum mismatch */
#define VMIO_ERR_UNSUPPORTED (-8) /* > 8 channels or > 32 bits per sample */
#define VMIO_ERR_CAPACITY (-9)    /* output buffer too small for the decoded stream */
#define VMIO_ERR_RESIDUAL (-10)   /* residual partitioning inconsistent with the block size / predictor order */
#define VMIO_ERR_NOMEM (-11)      /* scratch allocation failed */
#define VMIO_ERR_IO (-12)         /* file could not be opened / read */
#define VMIO_ERR_SHORT (-13)      /* a file holds fewer samples than the caller's index promised */

typedef struct vmio_flac_info {
    uint32_t sample_rate;
    uint32_t channels;
    uint32_t bits_per_sample;
    uint32_t min_blocksize;
    uint32_t max_blocksize;
    uint64_t total_samples; /* per channel; 0 = not recorded in the stream */
    uint8_t md5[16];        /* MD5 of the unencoded little-endian interleaved PCM; all zero = not recorded */
} vmio_flac_info;

int vmio_version(void);
const char* vmio_error_string(int code);

/* Parse the stream marker and the STREAMINFO block of a FLAC stream held in memory. */
int vmio_flac_probe(const uint8_t* data, size_t len, vmio_flac_info* info);

/* Decode a whole FLAC stream held in memory.
 *   out_i32 / out_f64: either may be NULL; interleaved (frame-major) output of `capacity_frames` * channels
 *                      elements.  out_f64 = sample / 2^(bits_per_sample-1) (the `soundfile.read` convention,
 *                      voicemap/librispeech.py:104).  With both NULL the stream is decoded and checked but not
 *                      stored, which is how a caller sizes its buffer when total_samples is 0.
 *   info:              optional, filled as by vmio_flac_probe.
 * Every frame's CRC-8 and CRC-16 is verified.  Returns the number of decoded frames (samples per channel) or a
 * negative VMIO_ERR_* code. */
int64_t vmio_flac_decode(const uint8_t* data, size_t len, int32_t* out_i32, double* out_f64,
                         uint64_t capacity_frames, vmio_flac_info* info);

/* Decode only samples [start, start + count) of each channel (the batcher wants a 3 s fragment of a ~12 s utterance,
 * voicemap/librispeech.py:105-112).  Frames carry no byte length, so the decoder jumps to a byte offset guessed from the
 * average compression ratio, resynchronises on the next frame that passes both CRCs, and reads its position from the
 * frame header; needs STREAMINFO's total_samples (otherwise it walks the stream from the start).  Output layout as
 * vmio_flac_decode, holding `count` frames.  Returns the number of frames written (< count when the stream ends
 * first) or a negative VMIO_ERR_* code. */
int64_t vmio_flac_decode_range(const uint8_t* data, size_t len, uint64_t start, uint64_t count, int32_t* out_i32,
                               double* out_f64, vmio_flac_info* info);

/* One batch of training clips in one call: row i of `out` (n rows of `want` doubles, the (B, T) array of
 * voicemap/librispeech.py:179-186 before its channel axis is added) becomes `lead[i]` zeros, samples
 * [start[i], start[i] + count[i]) of the mono file paths[i] scaled like vmio_flac_decode's f64 output, and zeros up to
 * `want` -- i.e. `__getitem__`'s crop and padding (voicemap/librispeech.py:105-124) with the random offsets chosen by the
 * caller.  `lead` may be NULL (no leading zeros).  Files are read and decoded on up to `threads` threads (the calling
 * thread included), each decoding only the frames under its fragment.  Returns 0, or the first error with its row in
 * *failed_row (optional); VMIO_ERR_SHORT when a file ends before start + count. */
int vmio_flac_read_fragments(const char* const* paths, const uint64_t* start, const uint64_t* count, const uint64_t* lead,
                             size_t n, double* out, size_t want, int threads, int64_t* failed_row);

/* Same as vmio_flac_decode for a file on disk (read fully into scratch memory first). */
int64_t vmio_flac_read_file(const char* path, int32_t* out_i32, double* out_f64, uint64_t capacity_frames,
                            vmio_flac_info* info);

/* STREAMINFO of a file on disk without decoding it: reads only the first bytes.  This is what indexing a corpus
 * needs (voicemap/librispeech.py:267-275 decodes every file just to learn its length). */
int vmio_flac_probe_file(const char* path, vmio_flac_info* info);

#ifdef __cplusplus
}
#endif
#endif /* VOICEMAP_IO_H */
