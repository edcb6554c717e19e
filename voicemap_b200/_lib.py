"""ctypes binding of libvoicemap_b200.so (the C ABI declared in include/voicemap_b200.h).

There is no fallback: if the library is missing or a call fails, a ``VoicemapB200Error`` is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvoicemap_b200.so")

VM_OK = 0
VM_ERR_SHAPE = -1
VM_ERR_UNSUPPORTED = -2
VM_ERR_CUDA = -3
VM_ERR_ARCH = -4
VM_METRIC_UNIFORM_EUCLIDEAN = 0
VM_METRIC_WEIGHTED_L1 = 1
VM_LOSS_NONE = 0
VM_LOSS_CONTRASTIVE = 1
VM_LOSS_BCE = 2

_vp, _i, _f, _sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t

# name -> (restype, argtypes); kept in one table so tests can check it against the header.
SIGNATURES = {
    "vm_version": (_i, []),
    "vm_last_error_string": (C.c_char_p, []),
    "vm_check_device": (_i, []),
    "vm_conv1_wpack_bytes": (_sz, [_i]),
    "vm_conv3_wpack_bytes": (_sz, [_i, _i]),
    "vm_epi_bytes": (_sz, [_i]),
    "vm_conv3_num_position_tiles": (_i, [_i]),
    "vm_padded_channels": (_i, [_i]),
    "vm_pack_conv1": (_i, [_vp] * 6 + [_f, _i, _vp, _vp, _vp]),
    "vm_pack_conv3": (_i, [_vp] * 6 + [_f, _i, _i, _vp, _vp, _vp]),
    "vm_conv1_relu_bn_pool_fwd": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _vp]),
    "vm_conv3_relu_bn_pool2_fwd": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "vm_gmax_dense_fwd": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "vm_pair_head_loss_fwd": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp]),
    "vm_nshot_score": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "vm_split_planes": (_i, [_vp, _sz, _vp, _vp, _vp]),
    "vm_merge_planes": (_i, [_vp, _vp, _sz, _vp, _vp]),
    "vm_split_planes_q": (_i, [_vp, _sz, _vp, _vp, _vp]),
    "vm_merge_planes_q": (_i, [_vp, _vp, _sz, _vp, _vp]),
    "vm_encoder_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "vm_encoder_fwd": (_i, [_vp, _i, _i, _i, _i, C.POINTER(_vp), C.POINTER(_vp), _vp, _vp, _i, _vp, _vp, _i, _vp]),
    "vm_set_option": (_i, [C.c_char_p, _i]),
    "vm_preprocess_scratch_bytes": (_sz, [_i]),
    "vm_preprocess_stats": (_i, [_vp, _i, _i, _i, _i, _f, _vp, _vp, _vp]),
    "vm_encoder_fwd_raw": (_i, [_vp, _i, _i, _i, _i, _f, _i, _i, C.POINTER(_vp), C.POINTER(_vp), _vp, _vp, _i, _vp, _vp,
                                _i, _vp]),
    # training
    "vm_pack_conv1_raw": (_i, [_vp, _vp, _i, _vp, _vp, _vp]),
    "vm_pack_conv3_raw": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp]),
    "vm_pack_conv3_dgrad": (_i, [_vp, _i, _i, _vp, _vp, _vp]),
    "vm_pack_train": (_i, [C.POINTER(_vp), C.POINTER(_vp), _i, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp),
                           C.POINTER(_vp), _vp]),
    "vm_conv1_train_fwd": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "vm_conv3_train_fwd": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "vm_conv3_dgrad": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "vm_stat_rows_per_clip": (_i, [_i]),
    "vm_conv3_train_rows_per_clip": (_i, [_i]),
    "vm_bn_stats_finalize": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _f, _f, _vp, _vp, _vp, _vp, _vp]),
    "vm_reduce_scratch_bytes": (_sz, [_i, _i]),
    "vm_bn_pool_fwd": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "vm_bn_gmax_fwd": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "vm_dense_fwd": (_i, [_vp, _i, _i, _vp, _vp, _i, _vp, _vp]),
    "vm_siamese_head_train": (_i, [_vp, _i, _i, _i, _vp, _vp, _i, _vp, _vp, _vp, _i, _f, _vp, _vp, _vp, _vp, _vp, _vp,
                                   _vp, _vp, _vp, _vp, _vp]),
    "vm_pair_head_loss_bwd": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _i, _f, _vp, _vp, _vp, _vp, _vp]),
    "vm_dense_bwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "vm_bn_bwd_scratch_elems": (_sz, [_i]),
    "vm_bn_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                       _vp, _vp, _i, _vp]),
    "vm_bn_stats_sums": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "vm_bn_stats_from_sums": (_i, [_vp, C.c_double, _i, _i, _vp, _vp, _f, _f, _vp, _vp, _vp, _vp]),
    "vm_bn_bwd_sums": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "vm_bn_bwd_from_sums": (_i, [_vp, _vp, C.c_double, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp,
                                 _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "vm_p2p_buffer_bytes": (_sz, []),
    "vm_p2p_alloc": (_i, [C.POINTER(_vp)]),
    "vm_p2p_free": (_i, [_vp]),
    "vm_p2p_export": (_i, [_vp, C.c_char_p]),
    "vm_p2p_import": (_i, [C.c_char_p, C.POINTER(_vp)]),
    "vm_p2p_unimport": (_i, [_vp]),
    "vm_bn_stats_finalize_peers": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _f, _f, _vp, _vp, _vp, _vp, C.POINTER(_vp), _i, _i,
                                        C.c_uint32, C.c_double, _vp, _vp, _vp]),
    "vm_bn_bwd_peers": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                             _vp, _vp, _vp, _i, C.POINTER(_vp), _i, _i, C.c_uint32, C.c_double, _vp, _vp, _vp]),
    "vm_wgrad3": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _sz, _vp, _vp]),
    "vm_wgrad1": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _sz, _vp, _vp]),
    "vm_adam_step": (_i, [_vp, _vp, _vp, _vp, _sz, _vp, _f, _f, _f, _f, _f, _f, _vp]),
}


class VoicemapB200Error(RuntimeError):
    pass


_lib = None


def load():
    """Load the shared library once and attach argument types.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VoicemapB200Error(
            f"{LIB_PATH} not found: build it with `python -m voicemap_b200.build` "
            "(voicemap_b200 has no CPU or PyTorch fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    return load().vm_last_error_string().decode("utf-8", "replace")


def check(rc: int, what: str):
    if rc != VM_OK:
        raise VoicemapB200Error(f"{what} failed with code {rc}: {last_error()}")
