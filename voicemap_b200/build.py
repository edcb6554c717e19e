"""In-tree build of libvoicemap_b200.so (hand-written sm_100a CUDA behind a C ABI) and of libvoicemap_io.so (the
host-side FLAC decoder of the batcher, plain C).

nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels to the GPU box with the
repo snapshot.  Usage: ``python -m voicemap_b200.build [--force]``.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libvoicemap_b200.so")
IO_LIB_PATH = os.path.join(HERE, "libvoicemap_io.so")
IO_SOURCES = ["vm_flac.c"]
IO_HEADERS = [os.path.join("..", "..", "include", "voicemap_io.h")]
CC_FLAGS = ["-O3", "-std=c99", "-Wall", "-Wextra", "-fwrapv", "-fPIC", "-shared"]  # -fwrapv: hostile streams may overflow the predictor
SOURCES = ["vm_api.cu", "vm_conv1.cu", "vm_conv3.cu", "vm_head.cu", "vm_train.cu", "vm_wgrad.cu"]
HEADERS = ["vm_common.cuh", "vm_kernels.h", "vm_p2p.cuh", os.path.join("..", "..", "include", "voicemap_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]


def _nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libvoicemap_b200.so")
    return nvcc


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_io_library(force: bool = False) -> str:
    """gcc build of the audio decoder (no CUDA involved)."""
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        raise RuntimeError("gcc not found: cannot build libvoicemap_io.so")
    srcs = [os.path.join(CSRC, s) for s in IO_SOURCES]
    if force or _stale(IO_LIB_PATH, srcs + [os.path.join(CSRC, h) for h in IO_HEADERS]):
        r = subprocess.run([cc, *CC_FLAGS, "-o", IO_LIB_PATH, *srcs, "-lpthread"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("gcc failed:\n" + r.stdout + r.stderr)
    return IO_LIB_PATH


def build_library(force: bool = False, verbose: bool = False) -> str:
    build_io_library(force)
    nvcc = _nvcc()
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    objs = []
    jobs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(CSRC, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            jobs.append([nvcc, *NVCC_FLAGS, "-c", s, "-o", o])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        return r.stderr

    if jobs:
        with ThreadPoolExecutor(max_workers=4) as ex:
            for log in ex.map(run, jobs):
                if verbose:
                    print(log)
    if force or jobs or _stale(LIB_PATH, objs):
        run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH, *objs])
    return LIB_PATH


if __name__ == "__main__":
    path = build_library(force="--force" in sys.argv, verbose=True)
    print("built", path)
