"""The C-ABI library loads and exports every symbol include/voicemap_b200.h declares (no compute calls: CPU box)."""
import ctypes
import os
import re

from voicemap_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "voicemap_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vm_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(built_library):
    lib = ctypes.CDLL(built_library)
    names = _declared_symbols()
    assert len(names) >= 15
    for name in names:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"


def test_ctypes_table_matches_header(built_library):
    assert sorted(_lib.SIGNATURES) == _declared_symbols()
    lib = _lib.load()
    assert lib.vm_version() == 100


def test_size_helpers(built_library):
    lib = _lib.load()
    assert lib.vm_padded_channels(128) == 128 and lib.vm_padded_channels(192) == 256
    assert lib.vm_conv1_wpack_bytes(128) == 16384
    assert lib.vm_conv3_wpack_bytes(128, 256) == 3 * 3 * 256 * 128 * 2   # planes: fp16 hi, fp16 lo, e5m2x2 Q
    assert lib.vm_epi_bytes(384) == 384 * 16
    assert lib.vm_conv3_num_position_tiles(3000) == 24
    # workspace: planes of the three stored activations + gmax partials (fp16 hi+lo = 4 bytes/element)
    n, l, f = 8, 12000, 128
    expect = 4 * n * (3000 * f + 1500 * 2 * f + 750 * 3 * f) + 4 * n * 6 * 4 * f
    got = lib.vm_encoder_workspace_bytes(n, l, f, 4)
    assert expect <= got <= expect + 8 * 1024
    assert lib.vm_encoder_workspace_bytes(8, 16, 128, 4) == 0  # too short for four pooling stages
    assert lib.vm_encoder_workspace_bytes(8, 16, 128, 2) > 0   # older architecture: pools 2*2*2*2
    assert lib.vm_encoder_workspace_bytes(8, 4096, 128, 3) == 0


def test_errors_are_codes_not_exceptions(built_library):
    lib = _lib.load()
    assert lib.vm_set_option(b"no_such_option", 1) == _lib.VM_ERR_SHAPE
    assert b"unknown key" in lib.vm_last_error_string()
    # null pointers are rejected before any CUDA call is made
    rc = lib.vm_conv1_relu_bn_pool_fwd(None, 1, 1024, 128, 4, None, None, None, None, 3, None)
    assert rc == _lib.VM_ERR_SHAPE


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    try:
        _lib.load()
    except _lib.VoicemapB200Error as exc:
        assert "no CPU or PyTorch fallback" in str(exc)
    else:
        raise AssertionError("load() must raise when the CUDA library is missing")


def test_headers_are_plain_c_and_a_c_program_links(built_library, tmp_path):
    """The boundary is a C ABI: both headers compile as strict C99 (no C++-isms, no torch types), and a C program linked
    against the two libraries calls their host-only entry points."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        import pytest
        pytest.skip("no gcc")
    from voicemap_b200.build import IO_LIB_PATH, build_io_library
    build_io_library()
    source = tmp_path / "client.c"
    source.write_text(r'''
#include <stdio.h>
#include "voicemap_b200.h"
#include "voicemap_io.h"
int main(void) {
    vmio_flac_info info;
    const unsigned char junk[8] = {'R', 'I', 'F', 'F', 0, 0, 0, 0};
    int rc = vmio_flac_probe(junk, sizeof junk, &info);
    printf("%d %d %d %s|%d %lu %d\n", vm_version(), vmio_version(), rc, vmio_error_string(rc),
           vm_padded_channels(192), (unsigned long)vm_conv1_wpack_bytes(128), vm_conv3_num_position_tiles(3000));
    return 0;
}
''')
    include = os.path.join(ROOT, "include")
    libdir = os.path.dirname(built_library)
    syntax = subprocess.run([gcc, "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-I", include,
                             str(source)], capture_output=True, text=True)
    assert syntax.returncode == 0, syntax.stderr
    exe = tmp_path / "client"
    link = subprocess.run([gcc, "-std=c99", "-I", include, str(source), "-o", str(exe), built_library, IO_LIB_PATH,
                           "-Wl,-rpath," + libdir], capture_output=True, text=True)
    assert link.returncode == 0, link.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True)
    assert run.returncode == 0, run.stderr
    assert run.stdout.strip() == "100 100 -2 not a FLAC stream (no fLaC marker / STREAMINFO)|256 16384 24"
