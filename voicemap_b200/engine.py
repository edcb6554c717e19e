"""Device-side engine of the voicemap encoder: owns the weights (Keras layout, fp32, on the GPU), their packed
tensor-core form, the activation workspace, and enqueues the C-ABI kernels on torch's current CUDA stream.

PyTorch is used for device memory, streams and (later) torch.distributed only -- every FLOP of the path runs in
libvoicemap_b200.so.  Reference semantics: voicemap/models.py:6-81.
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict

import numpy as np
import torch

from . import _lib

BN_EPS = 1e-3  # keras.layers.BatchNormalization default
POOLS = (4, 2, 2, 2)
PRECISION_FP32_GRADE = 3   # fp16 (hi, lo) planes, three MMAs per K step
PRECISION_THROUGHPUT = 1   # fp16 hi plane only
PRECISION_MIXED = 2        # fp16 main product + one fp8 (e5m2 pairs) product for both corrections: 8 MMAs per K chunk


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def encoder_param_shapes(filters, embedding_dimension):
    f = filters
    shapes = OrderedDict()
    cin = 1
    for i, (k, mult) in enumerate(((32, 1), (3, 2), (3, 3), (3, 4)), start=1):
        cout = mult * f
        shapes[f"conv{i}_kernel"] = (k, cin, cout)
        shapes[f"conv{i}_bias"] = (cout,)
        shapes[f"bn{i}_gamma"] = (cout,)
        shapes[f"bn{i}_beta"] = (cout,)
        shapes[f"bn{i}_mean"] = (cout,)
        shapes[f"bn{i}_var"] = (cout,)
        cin = cout
    shapes["dense_kernel"] = (4 * f, embedding_dimension)
    shapes["dense_bias"] = (embedding_dimension,)
    return shapes


class EncoderEngine:
    """get_baseline_convolutional_encoder (voicemap/models.py:6-41) in eval mode on one B200."""

    def __init__(self, filters, embedding_dimension, device=None, precision=PRECISION_FP32_GRADE, first_pool=4):
        """first_pool: size of the first MaxPool1D -- 4 (voicemap/models.py:19) or 2 (the older architecture of the
        checkpoint the reference ships under models/n_seconds/)."""
        if first_pool not in (2, 4):
            raise ValueError("first_pool must be 4 (voicemap/models.py:19) or 2 (older checkpoints)")
        self.first_pool = int(first_pool)
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.VoicemapB200Error("voicemap_b200 needs a CUDA device (no CPU fallback)")
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        with torch.cuda.device(self.device):
            _lib.check(self.lib.vm_check_device(), "vm_check_device")
        self.filters = int(filters)
        self.embedding_dimension = int(embedding_dimension)
        self.precision = int(precision)
        self.channels = [self.filters * m for m in (1, 2, 3, 4)]
        self.params = OrderedDict()
        for name, shape in encoder_param_shapes(filters, embedding_dimension).items():
            self.params[name] = torch.zeros(shape, dtype=torch.float32, device=self.device)
        self.wpack = []
        self.epi = []
        cin = 1
        for i, cout in enumerate(self.channels):
            nbytes = (self.lib.vm_conv1_wpack_bytes(cout) if i == 0 else self.lib.vm_conv3_wpack_bytes(cin, cout))
            self.wpack.append(torch.zeros(nbytes, dtype=torch.uint8, device=self.device))
            self.epi.append(torch.zeros(self.lib.vm_epi_bytes(cout) // 4, dtype=torch.float32, device=self.device))
            cin = cout
        self._packed = False
        self._workspace = None
        self._ws_key = None

    # ------------------------------------------------------------------ weights
    def set_weights(self, params):
        """params: mapping name -> array in Keras layout (see encoder_param_shapes)."""
        for name, dst in self.params.items():
            src = torch.as_tensor(np.ascontiguousarray(np.asarray(params[name], dtype=np.float32)))
            if tuple(src.shape) != tuple(dst.shape):
                raise ValueError(f"{name}: expected shape {tuple(dst.shape)}, got {tuple(src.shape)}")
            dst.copy_(src)
        self._packed = False

    def get_weights(self):
        return OrderedDict((k, v.detach().cpu().numpy()) for k, v in self.params.items())

    def pack(self):
        """Fold BN into per-channel constants and write the fp16 (hi, lo) tensor-core weight planes."""
        p = self.params
        with torch.cuda.device(self.device):
            cin = 1
            for i, cout in enumerate(self.channels, start=1):
                args = [_ptr(p[f"conv{i}_kernel"]), _ptr(p[f"conv{i}_bias"]), _ptr(p[f"bn{i}_gamma"]),
                        _ptr(p[f"bn{i}_beta"]), _ptr(p[f"bn{i}_mean"]), _ptr(p[f"bn{i}_var"]), C.c_float(BN_EPS)]
                if i == 1:
                    rc = self.lib.vm_pack_conv1(*args, cout, _ptr(self.wpack[0]), _ptr(self.epi[0]), _stream())
                else:
                    rc = self.lib.vm_pack_conv3(*args, cin, cout, _ptr(self.wpack[i - 1]), _ptr(self.epi[i - 1]),
                                                _stream())
                _lib.check(rc, f"vm_pack_conv{1 if i == 1 else 3} (block {i})")
                cin = cout
        self._packed = True

    # ------------------------------------------------------------------ forward
    def _get_workspace(self, n, length):
        key = (n, length)
        if self._ws_key != key:
            nbytes = self.lib.vm_encoder_workspace_bytes(n, length, self.filters, self.first_pool)
            nbytes += self.lib.vm_preprocess_scratch_bytes(n) if nbytes else 0
            if nbytes == 0:
                raise ValueError(f"input of shape ({n}, {length}) is too short for the encoder (L >= {8 * self.first_pool})")
            if self._workspace is None or self._workspace.numel() < nbytes + 1024:   # grow-only
                self._workspace = None
                self._workspace = torch.empty(nbytes + 1024, dtype=torch.uint8, device=self.device)
            self._ws_key = key
        base = self._workspace.data_ptr()
        return C.c_void_p((base + 1023) // 1024 * 1024)

    def forward(self, x, out=None, precision=None):
        """x: contiguous fp32 tensor (N, L) or (N, L, 1), on the device or in PINNED host memory (block 1 then reads
        the waveform over PCIe/C2C directly -- pinned allocations are device-mapped under unified addressing; the
        caller keeps the tensor alive and unchanged until the stream has run).  Returns (N, E) fp32 (CUDA).
        ``precision`` overrides the engine's arithmetic for this call (the packed weights carry the planes of every
        mode and the workspace layout is the same)."""
        if x.dim() == 3:
            if x.shape[2] != 1:
                raise ValueError("encoder input must have one channel: (N, L, 1)")
            x = x.reshape(x.shape[0], x.shape[1])
        if x.dtype != torch.float32 or not (x.is_cuda or x.is_pinned()) or not x.is_contiguous():
            raise ValueError("encoder input must be a contiguous float32 tensor on the device or in pinned memory")
        if not self._packed:
            self.pack()
        n, length = x.shape
        if out is None:
            out = torch.empty((n, self.embedding_dimension), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            ws = self._get_workspace(n, length)
            wp = (C.c_void_p * 4)(*[t.data_ptr() for t in self.wpack])
            ep = (C.c_void_p * 4)(*[t.data_ptr() for t in self.epi])
            rc = self.lib.vm_encoder_fwd(_ptr(x), n, length, self.filters, self.first_pool, wp, ep, _ptr(self.params["dense_kernel"]),
                                         _ptr(self.params["dense_bias"]), self.embedding_dimension, ws, _ptr(out),
                                         int(precision) if precision is not None else self.precision, _stream())
        _lib.check(rc, "vm_encoder_fwd")
        return out

    def forward_raw(self, x, downsampling=4, whiten_groups=1, rms=0.038021, out=None):
        """Raw audio (N, T) fp32 on the device -> embeddings, with the reference's preprocessing
        (voicemap/utils.py:22-34: x[:, ::downsampling], whiten) fused into block 1.  ``whiten_groups`` = number of
        separate whiten() calls the batch stands for (their scales are batch-global per call); 0 = no whitening."""
        if x.dim() == 3:
            x = x.reshape(x.shape[0], x.shape[1])
        if x.dtype != torch.float32 or not x.is_cuda or not x.is_contiguous():
            raise ValueError("raw input must be a contiguous CUDA float32 tensor")
        if not self._packed:
            self.pack()
        n, t = x.shape
        length = (t + downsampling - 1) // downsampling
        if out is None:
            out = torch.empty((n, self.embedding_dimension), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            ws = self._get_workspace(n, length)
            wp = (C.c_void_p * 4)(*[w.data_ptr() for w in self.wpack])
            ep = (C.c_void_p * 4)(*[e.data_ptr() for e in self.epi])
            rc = self.lib.vm_encoder_fwd_raw(_ptr(x), n, t, int(downsampling), int(whiten_groups), C.c_float(rms),
                                             self.filters, self.first_pool, wp, ep,
                                             _ptr(self.params["dense_kernel"]), _ptr(self.params["dense_bias"]),
                                             self.embedding_dimension, ws, _ptr(out), self.precision, _stream())
        _lib.check(rc, "vm_encoder_fwd_raw")
        return out

    # ------------------------------------------------------------------ per-block views (tests, profiling)
    def split_planes(self, x):
        """fp32 -> the plane pair of this engine's precision: (hi, lo) fp16, or (hi, Q) with precision 2."""
        hi = torch.empty(x.shape, dtype=torch.float16, device=self.device)
        lo = torch.empty(x.shape, dtype=torch.float16, device=self.device)
        fn = self.lib.vm_split_planes_q if self.precision == PRECISION_MIXED else self.lib.vm_split_planes
        _lib.check(fn(_ptr(x), x.numel(), _ptr(hi), _ptr(lo), _stream()), "vm_split_planes")
        return hi, lo

    def merge_planes(self, hi, lo):
        x = torch.empty(hi.shape, dtype=torch.float32, device=self.device)
        fn = self.lib.vm_merge_planes_q if self.precision == PRECISION_MIXED else self.lib.vm_merge_planes
        _lib.check(fn(_ptr(hi), _ptr(lo), hi.numel(), _ptr(x), _stream()), "vm_merge_planes")
        return x

    def block1(self, x, out=None):
        """x (N, L) fp32 -> planes (N, L//first_pool, f).  ``out`` = (hi, lo) reuses caller-owned planes."""
        if not self._packed:
            self.pack()
        n, length = x.shape
        f = self.channels[0]
        if out is not None:
            hi, lo = out
            assert hi.shape == (n, length // self.first_pool, f) and lo.shape == hi.shape and hi.dtype == torch.float16
        else:
            hi = torch.empty((n, length // self.first_pool, f), dtype=torch.float16, device=self.device)
            lo = torch.empty_like(hi)
        rc = self.lib.vm_conv1_relu_bn_pool_fwd(_ptr(x), n, length, f, self.first_pool, _ptr(self.wpack[0]),
                                                _ptr(self.epi[0]), _ptr(hi), _ptr(lo), self.precision, _stream())
        _lib.check(rc, "vm_conv1_relu_bn_pool_fwd")
        return hi, lo

    def block3(self, index, in_hi, in_lo, gmax=False, out=None):
        """Block `index` in {2,3,4}: planes (N, L, Cin) -> planes (N, L//2, Cout), or gmax partials.
        ``out`` = (hi, lo) planes, or the partials tensor with gmax=True, reuses caller-owned buffers."""
        if not self._packed:
            self.pack()
        n, length, cin = in_hi.shape
        cout = self.channels[index - 1]
        assert cin == self.channels[index - 2]
        if gmax:
            t = self.lib.vm_conv3_num_position_tiles(length)
            cpad = self.lib.vm_padded_channels(cout)
            part = out if out is not None else torch.empty((n, t, cpad), dtype=torch.float32, device=self.device)
            assert part.shape == (n, t, cpad) and part.dtype == torch.float32
            rc = self.lib.vm_conv3_relu_bn_pool2_fwd(_ptr(in_hi), _ptr(in_lo), n, length, cin, cout,
                                                     _ptr(self.wpack[index - 1]), _ptr(self.epi[index - 1]),
                                                     None, None, _ptr(part), self.precision, _stream())
            _lib.check(rc, "vm_conv3_relu_bn_pool2_fwd(gmax)")
            return part
        if out is not None:
            hi, lo = out
            assert hi.shape == (n, length // 2, cout) and lo.shape == hi.shape and hi.dtype == torch.float16
        else:
            hi = torch.empty((n, length // 2, cout), dtype=torch.float16, device=self.device)
            lo = torch.empty_like(hi)
        rc = self.lib.vm_conv3_relu_bn_pool2_fwd(_ptr(in_hi), _ptr(in_lo), n, length, cin, cout,
                                                 _ptr(self.wpack[index - 1]), _ptr(self.epi[index - 1]), _ptr(hi),
                                                 _ptr(lo), None, self.precision, _stream())
        _lib.check(rc, "vm_conv3_relu_bn_pool2_fwd")
        return hi, lo

    def gmax_dense(self, part, with_gmax=False):
        n, t, cpad = part.shape
        c = self.channels[3]
        emb = torch.empty((n, self.embedding_dimension), dtype=torch.float32, device=self.device)
        g = torch.empty((n, c), dtype=torch.float32, device=self.device) if with_gmax else None
        rc = self.lib.vm_gmax_dense_fwd(_ptr(part), n, t, c, _ptr(self.epi[3]), _ptr(self.params["dense_kernel"]),
                                        _ptr(self.params["dense_bias"]), self.embedding_dimension, _ptr(g),
                                        _ptr(emb), _stream())
        _lib.check(rc, "vm_gmax_dense_fwd")
        return (emb, g) if with_gmax else emb


def pair_head_loss(e1, e2, head_w, head_b, metric="uniform_euclidean", y_true=None, loss=None):
    """Siamese head on embeddings (voicemap/models.py:55-69) + optional loss.  Returns (prob, dist, loss)."""
    lib = _lib.load()
    metric_id = {"uniform_euclidean": _lib.VM_METRIC_UNIFORM_EUCLIDEAN, "weighted_l1": _lib.VM_METRIC_WEIGHTED_L1}
    if metric not in metric_id:
        raise NotImplementedError(metric)
    loss_id = {None: _lib.VM_LOSS_NONE, "contrastive": _lib.VM_LOSS_CONTRASTIVE,
               "binary_crossentropy": _lib.VM_LOSS_BCE}[loss]
    n, e = e1.shape
    dev = e1.device
    prob = torch.empty((n, 1), dtype=torch.float32, device=dev)
    dist = torch.empty((n, 1), dtype=torch.float32, device=dev) if metric == "uniform_euclidean" else None
    lossv = torch.zeros((1,), dtype=torch.float32, device=dev) if loss is not None else None
    with torch.cuda.device(dev):
        rc = lib.vm_pair_head_loss_fwd(_ptr(e1), _ptr(e2), n, e, metric_id[metric], _ptr(head_w), _ptr(head_b),
                                       _ptr(y_true), loss_id, _ptr(dist), _ptr(prob), _ptr(lossv), _stream())
    _lib.check(rc, "vm_pair_head_loss_fwd")
    return prob, dist, lossv
