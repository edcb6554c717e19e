/* Mutation fuzzer for the FLAC decoder (voicemap_b200/csrc/vm_flac.c), meant to be built with
 * -fsanitize=address,undefined: loads seed streams, damages them (bit flips, byte stores, truncation, block copies)
 * and calls every entry point.  The decoder must return -- an error code or a count -- never crash, hang, read or
 * write out of bounds.  usage: fuzz_flac <iterations> <seed> <file.flac>...     (tools/fuzz_flac.sh builds and runs it) */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/voicemap_io.h"

static uint64_t rng_state;
static uint32_t rnd(void) {
    rng_state ^= rng_state << 13;
    rng_state ^= rng_state >> 7;
    rng_state ^= rng_state << 17;
    return (uint32_t)(rng_state >> 16);
}

int main(int argc, char** argv) {
    if (argc < 4) return 2;
    const long iterations = atol(argv[1]);
    rng_state = strtoull(argv[2], NULL, 10) * 2654435761u + 88172645463325252ull;
    const int nseeds = argc - 3;
    uint8_t** seeds = malloc(sizeof(*seeds) * nseeds);
    size_t* sizes = malloc(sizeof(*sizes) * nseeds);
    for (int i = 0; i < nseeds; ++i) {
        FILE* f = fopen(argv[3 + i], "rb");
        if (!f) { perror(argv[3 + i]); return 2; }
        fseek(f, 0, SEEK_END);
        sizes[i] = (size_t)ftell(f);
        fseek(f, 0, SEEK_SET);
        seeds[i] = malloc(sizes[i]);
        if (fread(seeds[i], 1, sizes[i], f) != sizes[i]) return 2;
        fclose(f);
    }
    long ok = 0, failed = 0;
    long histogram[16] = {0};
    for (long it = 0; it < iterations; ++it) {
        const int which = (int)(rnd() % (uint32_t)nseeds);
        size_t len = sizes[which];
        /* exact-size heap copy: any read past the end is an ASan report */
        uint8_t* data = malloc(len ? len : 1);
        memcpy(data, seeds[which], len);
        const int edits = 1 + (int)(rnd() % 4);
        for (int e = 0; e < edits && len > 0; ++e) {
            const uint32_t kind = rnd() % 6;
            const size_t at = rnd() % len;
            if (kind == 0) data[at] ^= (uint8_t)(1u << (rnd() % 8));
            else if (kind == 1) data[at] = (uint8_t)rnd();
            else if (kind == 2) data[at] = (rnd() & 1) ? 0xff : 0x00;
            else if (kind == 3 && len > 8) len = 1 + rnd() % len; /* truncate (the tail stays allocated but unused) */
            else if (kind == 4) { const size_t from = rnd() % len, n = 1 + rnd() % 64;
                                  for (size_t k = 0; k < n && at + k < len && from + k < len; ++k) data[at + k] = data[from + k]; }
            else if (at < 64) data[at] = (uint8_t)rnd(); /* concentrate on the headers */
        }
        uint8_t* exact = malloc(len ? len : 1);
        memcpy(exact, data, len);
        free(data);

        vmio_flac_info info;
        vmio_flac_probe(exact, len, &info);
        const int64_t frames = vmio_flac_decode(exact, len, NULL, NULL, 0, &info);
        if (frames >= 0) {
            ++ok;
            const uint64_t cap = (uint64_t)frames;
            int32_t* pcm = malloc(sizeof(int32_t) * (cap * info.channels + 1));
            double* f64 = malloc(sizeof(double) * (cap * info.channels + 1));
            const int64_t again = vmio_flac_decode(exact, len, pcm, f64, cap, NULL);
            if (again != frames) { fprintf(stderr, "iteration %ld: count pass %lld, store pass %lld\n", it, (long long)frames, (long long)again); return 1; }
            if (cap > 1 && vmio_flac_decode(exact, len, pcm, NULL, cap - 1, NULL) != VMIO_ERR_CAPACITY) {
                fprintf(stderr, "iteration %ld: short buffer accepted\n", it); return 1; }
            free(pcm);
            free(f64);
        } else {
            ++failed;
            histogram[(-frames) & 15]++;
        }
        for (int r = 0; r < 2; ++r) { /* partial decoding: whatever it returns, it must stay inside `count` frames */
            const uint64_t count = rnd() % 5000;
            const uint64_t start = rnd() % 60000;
            const uint32_t ch = (frames >= 0 || info.channels) && info.channels <= 8 ? (info.channels ? info.channels : 1) : 8;
            double* out = malloc(sizeof(double) * (count * ch + 1));
            const int64_t got = vmio_flac_decode_range(exact, len, start, count, NULL, out, NULL);
            if (got > (int64_t)count) { fprintf(stderr, "iteration %ld: range returned %lld > %llu\n", it, (long long)got, (unsigned long long)count); return 1; }
            free(out);
        }
        free(exact);
    }
    printf("%ld iterations: %ld decoded, %ld rejected; by error code:", iterations, ok, failed);
    for (int i = 1; i < 13; ++i) printf(" -%d:%ld", i, histogram[i]);
    printf("\n");
    for (int i = 0; i < nseeds; ++i) free(seeds[i]);
    free(seeds);
    free(sizes);
    return 0;
}
