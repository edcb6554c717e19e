#!/bin/bash
# Builds the FLAC decoder with AddressSanitizer + UBSan and fuzzes it with mutated copies of valid streams.
# usage: tools/fuzz_flac.sh [iterations per seed set]      (CPU only; ~1 minute for the default)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
WORK=${TMPDIR:-/tmp}/vm_fuzz_flac
mkdir -p "$WORK"
python - "$WORK" <<'PY'
import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
root = sys.argv[1]
from flac_writer import encode_flac, encode_flac_quick, id3v2_tag
rng = np.random.default_rng(0)
t = np.arange(3000)
mono = np.round(9000 * np.sin(t * 0.02) + 300 * rng.standard_normal(3000)).astype(np.int64)
stereo = np.stack([mono, np.round(0.7 * mono + 100 * rng.standard_normal(3000)).astype(np.int64)], axis=1)
seeds = {
    "lpc.flac": encode_flac(mono, 16000, 16, 1024, kind="lpc", lpc_order=8, partition_order=2),
    "fixed_escape.flac": encode_flac(mono[:2048], 16000, 16, 512, kind="fixed", fixed_order=3, partition_order=2, escape_partitions=(1,), method=1),
    "midside24.flac": encode_flac(stereo[:1152] * 200, 44100, 24, 576, kind="lpc", lpc_order=4, stereo="mid_side"),
    "variable.flac": encode_flac(mono[:1808], 16000, 16, kind="fixed", blocksizes=[192, 576, 1000, 40], record_length=False, id3v2=id3v2_tag(40), id3v1=True),
    "verbatim.flac": encode_flac_quick(mono, 16000, 256),
    "constant.flac": encode_flac_quick(None, constant=(-5, 9000)),
}
for name, data in seeds.items():
    open(os.path.join(root, name), "wb").write(data)
PY
gcc -O1 -g -std=gnu99 -fsanitize=address,undefined -fno-sanitize-recover=undefined -fwrapv \
    -o "$WORK/fuzz_flac" "$ROOT/tools/fuzz_flac.c" "$ROOT/voicemap_b200/csrc/vm_flac.c"
export ASAN_OPTIONS=detect_leaks=1:abort_on_error=1 UBSAN_OPTIONS=print_stacktrace=1
for seed in 1 2 3 4; do
    timeout 600 "$WORK/fuzz_flac" "${1:-40000}" $seed "$WORK"/*.flac
done
