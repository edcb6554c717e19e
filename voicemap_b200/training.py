"""Training on B200: what ``model.fit_generator`` does in the reference (experiments/train_siamese.py:56-94,
train_classifier.py:114-150, siamese_contrastive_loss.py:70-100) -- train-mode forward (batch-statistics
BatchNorm per encoder application, SpatialDropout1D), backward, Keras Adam with global-norm clipping, the Keras
callback protocol -- with every FLOP of the encoder/head in libvoicemap_b200.so.  PyTorch supplies device memory,
streams and (multi-GPU) ``torch.distributed.all_reduce`` of the flat gradient buffer.

Data parallelism: each rank trains on its own shard of the batch; ONE flat fp32 gradient
all-reduce (sum) per step over NCCL, then the identical clip + Adam update on every rank.  BatchNorm statistics
are those of the GLOBAL batch (synchronised BatchNorm: eight small all-reduces of per-channel double sums per step), so
W ranks x N/W pairs reproduce the reference's single-device step on N pairs; ``model.sync_batchnorm = False`` opts out.

A step is a fixed sequence of ~75 kernel launches on buffers that keep their addresses, so the launches are recorded once
per step shape with their ctypes arguments prebuilt (``_Plan``) and batches reach the device through pinned staging
buffers: the host issues a step in ~0.4 ms and never waits for the device inside one.
"""
from __future__ import annotations

import ctypes as C
import os
from collections import OrderedDict

import numpy as np
import torch

from . import _lib
from .engine import BN_EPS, PRECISION_FP32_GRADE, EncoderEngine, _ptr, _stream

BN_MOMENTUM = 0.99
POOLS = (4, 2, 2, 2)


def _check(rc, what):
    _lib.check(rc, what)


class _Plan:
    """A recorded sequence of C-ABI launches for one step shape.  Every buffer a training step touches is allocated
    once per (batch, length) and keeps its address, so the ctypes argument objects are built once and a step replays
    ``fn(*args)`` per kernel instead of re-wrapping ~20 pointers and scalars per launch.  At 16 pairs per GPU (the
    128-pair batch of BASELINE config[2] over 8 ranks) building the arguments cost more host time than the kernels took
    on the device (tools/train_bench.py: host_issue_ms_per_step).  ``host`` entries are python callables (the BatchNorm
    all-reduces of the data-parallel path) that run in sequence with the launches."""

    def __init__(self):
        self.ops = []

    def launch(self, fn, what, *args):
        self.ops.append((fn, args, what))

    def host(self, fn):
        self.ops.append((None, fn, None))

    def run(self):
        for fn, args, what in self.ops:
            if fn is None:
                args()
            else:
                rc = fn(*args)
                if rc:
                    _lib.check(rc, what)

    def run_timed(self, timings):
        """The same launches with a CUDA event after every one (live timings of a step: ncu serialises the launches
        and flushes caches between them, which inflates the small kernels).  ``timings``: list that receives
        (what, event before, event after) triples."""
        before = torch.cuda.Event(enable_timing=True)
        before.record()
        for fn, args, what in self.ops:
            if fn is None:
                args()
                continue
            rc = fn(*args)
            if rc:
                _lib.check(rc, what)
            after = torch.cuda.Event(enable_timing=True)
            after.record()
            timings.append((what, before, after))
            before = after


class TrainEngine:
    """One training step of encoder (+ siamese head | classifier head) on the current CUDA device."""

    def __init__(self, model, optimizer, loss, loss_scale=1.0, precision=PRECISION_FP32_GRADE, bwd_precision=None,
                 seed=0):
        """precision: arithmetic of the train-mode forward convolutions of blocks 2-4 -- 3 = fp16 x 3, 2 = fp16 + fp8
        correction product; block 1 always runs fp16 x 3.  bwd_precision: 3 = two-plane fp16
        gradients, three MMAs per step; 2 = one-plane fp16 gradients (each dU element rounded to 11 significant bits,
        unbiased) against two-plane activations / weights, two MMAs; 1 = one plane each."""
        from .models import EncoderModel, SiameseModel
        self.lib = _lib.load()
        self.model = model
        self.optimizer = optimizer
        self.loss = loss
        self.loss_scale = float(loss_scale)
        self.precision = int(precision)
        # default backward arithmetic: one-plane fp16 gradients and operands (gradients within 2e-3 of fp64 autograd,
        # measured 8.4e-4 at 64 pairs x 12000); `model.train_bwd_precision = 3` selects the two-plane form (5e-4 / 2.1e-4)
        self.bwd_precision = int(bwd_precision if bwd_precision is not None
                                 else getattr(model, "train_bwd_precision", 1))
        if self.precision not in (1, 2, 3) or self.bwd_precision not in (1, 2, 3):
            raise ValueError("precision and bwd_precision must be 1, 2 or 3")
        # Synchronised BatchNorm: set_sync_bn(allreduce, world) makes the batch statistics those of
        # the GLOBAL batch, as on the reference's single device; None = per-rank statistics
        self.sync_allreduce = None
        self.sync_world = 1
        self.sync_peers = None     # parallel.PeerExchange: the BatchNorm sums cross the ranks inside the kernels (NVLink)
        self.grad_buckets = None   # parallel.GradientBuckets: per-block asynchronous gradient all-reduce (data parallel)
        # weight gradients on a side stream: wgrad of block b only feeds the optimizer, so it runs beside the BatchNorm /
        # ReLU backward pass of block b-1 (HBM-bound, no shared memory) instead of in front of it (VOICEMAP_WGRAD_STREAM=0
        # or overlap_wgrad = False: everything on one stream, as the per-launch timeline needs it)
        self.overlap_wgrad = os.environ.get("VOICEMAP_WGRAD_STREAM", "1") != "0"
        self._side = None
        if isinstance(model, SiameseModel):
            self.kind = "siamese"
            self.encoder_model = model.encoder
            if loss not in ("binary_crossentropy", "contrastive_loss"):
                raise NotImplementedError(f"siamese training with loss {loss!r}")
            head = model.head_weights
            self.head_names = ["head_kernel", "head_bias"]
        elif isinstance(model, EncoderModel):
            self.kind = "classifier"
            self.encoder_model = model
            if model._head is None or model._head["activation"] != "softmax" or loss != "categorical_crossentropy":
                raise NotImplementedError("classifier training needs a Dense(softmax) head and "
                                          "categorical_crossentropy")
            head = OrderedDict((k, model.weights[k]) for k in ("head_kernel", "head_bias"))
            self.head_names = ["head_kernel", "head_bias"]
        else:
            raise TypeError("unknown model type")
        enc = self.encoder_model
        self.eng: EncoderEngine = enc._get_engine()
        self.device = self.eng.device
        self.filters, self.emb = enc.filters, enc.embedding_dimension
        self.channels = [self.filters * m for m in (1, 2, 3, 4)]
        self.pools = (enc.first_pool,) + POOLS[1:]
        self.dropout = float(enc.dropout)
        self.gen = torch.Generator(device=self.device)
        self.gen.manual_seed(int(seed))

        # ---- flat trainable-parameter buffer in Keras weight order (moving statistics excluded)
        self.layout = OrderedDict()
        off = 0
        for name, w in enc.weights.items():
            if name.endswith("_mean") or name.endswith("_var") or name.startswith("head_"):
                continue
            self.layout[name] = (off, tuple(w.shape))
            off += int(np.prod(w.shape))
        for name in self.head_names:
            self.layout[name] = (off, tuple(head[name].shape))
            off += int(np.prod(head[name].shape))
        self.nparams = off
        dev = self.device
        self.flat = torch.zeros(off, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(off, dtype=torch.float32, device=dev)
        self.adam_m = torch.zeros(off, dtype=torch.float32, device=dev)
        self.adam_v = torch.zeros(off, dtype=torch.float32, device=dev)
        self.sumsq = torch.zeros(1, dtype=torch.float64, device=dev)
        self.p = OrderedDict()   # parameter views
        self.g = OrderedDict()   # gradient views
        for name, (o, shape) in self.layout.items():
            n = int(np.prod(shape))
            self.p[name] = self.flat[o:o + n].view(shape)
            self.g[name] = self.grad[o:o + n].view(shape)
            src = head[name] if name in self.head_names else enc.weights[name]
            self.p[name].copy_(torch.from_numpy(np.ascontiguousarray(src, dtype=np.float32)))
        # moving statistics stay in the eval engine's tensors; trainable views are shared with it so that
        # predict()/validation see the current weights without copies
        self.moving = OrderedDict((k, self.eng.params[k]) for k in self.eng.params if k.endswith(("_mean", "_var")))
        for name in self.eng.params:
            if name in self.p:
                self.eng.params[name] = self.p[name]
        self.eng._packed = False

        # ---- packed operands for train mode
        self.wraw, self.eraw, self.wdg, self.edg = [], [], [None], [None]
        cin = 1
        for i, cout in enumerate(self.channels):
            nb = self.lib.vm_conv1_wpack_bytes(cout) if i == 0 else self.lib.vm_conv3_wpack_bytes(cin, cout)
            self.wraw.append(torch.zeros(nb, dtype=torch.uint8, device=dev))
            self.eraw.append(torch.zeros(self.lib.vm_epi_bytes(cout) // 4, dtype=torch.float32, device=dev))
            if i > 0:
                self.wdg.append(torch.zeros(self.lib.vm_conv3_wpack_bytes(cout, cin), dtype=torch.uint8, device=dev))
                self.edg.append(torch.zeros(self.lib.vm_epi_bytes(cin) // 4, dtype=torch.float32, device=dev))
            cin = cout
        self._buf_key = None
        self._plans = {}
        self._pin = [None, None, None, None, None, None]   # pinned staging ring (two clip sides + labels per step, twice)
        self._pin_next = 0
        self.iterations = 0

    # ------------------------------------------------------------------ buffers
    def _buffers(self, nb, length, groups):
        key = (nb, length, groups)
        if self._buf_key == key:
            return
        self.U16 = self.EXT = self.X = self.XL = self.XQ = self.redp = self.dU = self.dX = self.scr1 = self.scr2 = None   # free before regrowing
        dev, f32, f16 = self.device, torch.float32, torch.float16
        ls = [length]
        for p in self.pools:
            ls.append(ls[-1] // p)
        if ls[4] < 1:
            raise ValueError("clips are too short for the encoder")
        self.ls = ls  # ls[b] = un-pooled length of block b+1's conv; ls[b + 1] = its pooled length
        c = self.channels
        # per block: encoded un-pooled activations (fp16 + arg-max flag) and the fp32 extreme of every pool window
        self.U16 = [torch.empty((nb, ls[b], c[b]), dtype=torch.int16, device=dev) for b in range(4)]
        self.EXT = [torch.empty((nb, ls[b + 1], c[b]), dtype=f32, device=dev) for b in range(4)]
        # pooled, normalised activations = inputs of the next block's conv: fp16 hi plane, fp16 residual plane (forward
        # precision 3; the weight gradient's second plane whenever bwd_precision >= 2) and / or e5m2x2 Q plane (forward
        # precision 2)
        want_lo = self.precision == 3 or (self.precision == 2 and self.bwd_precision >= 2)
        self.X = [torch.empty((nb, ls[b + 1], c[b]), dtype=f16, device=dev) for b in range(3)]
        self.XL = [torch.empty((nb, ls[b + 1], c[b]), dtype=f16, device=dev) if want_lo else None for b in range(3)]
        self.XQ = [torch.empty((nb, ls[b + 1], c[b]), dtype=f16, device=dev) if self.precision == 2 else None
                   for b in range(3)]
        rows = [self.lib.vm_stat_rows_per_clip(ls[0])] + [self.lib.vm_conv3_train_rows_per_clip(ls[b]) for b in (1, 2, 3)]
        self.stat = [torch.empty((nb * rows[b], self.lib.vm_padded_channels(c[b]), 2), dtype=f32, device=dev)
                     for b in range(4)]
        self.stat_rows = rows
        self.bnc = [torch.empty((groups, c[b], 4), dtype=f32, device=dev) for b in range(4)]
        self.bwc = [torch.empty((groups, c[b], 4), dtype=f32, device=dev) for b in range(4)]
        self.gmax = torch.empty((nb, c[3]), dtype=f32, device=dev)
        self.jstar = torch.empty((nb, c[3]), dtype=torch.int32, device=dev)
        self.embv = torch.empty((nb, self.emb), dtype=f32, device=dev)
        self.d_emb = torch.empty((nb, self.emb), dtype=f32, device=dev)
        self.d_gmax = torch.empty((nb, c[3]), dtype=f32, device=dev)
        max_u = max(ls[b] * c[b] for b in range(4))
        # gradient planes, two sets used alternately by the blocks: block b-2 may overwrite a set only after block b's
        # weight gradient (side stream) has read it
        self.dU = [torch.empty((2 if self.bwd_precision == 3 else 1, nb * max_u), dtype=f16, device=dev)
                   for _ in range(2)]
        self.gabs = torch.zeros((4,), dtype=torch.int32, device=dev)   # per block: bits of max |s * dy| -> gradient scale
        # blocks 1-3: partial rows of the BatchNorm-backward sums, written by the dgrad epilogue of the block above
        self.redp = [torch.empty((nb * self.lib.vm_conv3_train_rows_per_clip(ls[b + 1]), self.lib.vm_padded_channels(c[b]), 2),
                                 dtype=f32, device=dev) for b in range(3)]
        max_x = max(ls[b + 1] * c[b] for b in range(3))
        self.dX = torch.empty(nb * max_x, dtype=f32, device=dev)
        scr_elems = self.lib.vm_bn_bwd_scratch_elems(nb)
        self.scr2 = torch.empty((scr_elems, 2), dtype=f32, device=dev)
        self.scr1 = torch.empty((scr_elems,), dtype=f32, device=dev)
        self.red = torch.empty(self.lib.vm_reduce_scratch_bytes(groups, max(c)) // 8, dtype=torch.float64, device=dev)
        self.sums = torch.zeros((2, groups * max(c) * 2), dtype=torch.float64, device=dev)   # [local | global]
        # wgrad split partials (the launcher lowers its split count to fit) / wgrad1 per-CTA partials
        w1_bytes = nb * ((ls[0] + 1023) // 1024) * 32 * c[0] * 4
        self.wpart = torch.empty(max(64 << 20, w1_bytes) // 4, dtype=f32, device=dev)
        self.xin = torch.empty((nb, length), dtype=f32, device=dev)            # the step's clips (both branches)
        self.yin = torch.empty((max(1, nb // groups),), dtype=f32, device=dev)   # siamese pair labels
        self.maskbuf = [torch.empty((nb, c[b]), dtype=f32, device=dev) for b in range(4)]
        self.prob = torch.empty((max(1, nb // groups), 1), dtype=f32, device=dev)
        self.lossv = torch.zeros((2,), dtype=f32, device=dev)      # [loss, accuracy] of the step
        self.pairrec = torch.empty((max(1, nb // groups), 4), dtype=f32, device=dev)   # per pair {loss, hit, dL/dz, dist}
        self.masks = [None] * 4
        self._plans = {}
        self._buf_key = key

    def _plan_pack(self, plan, st):
        """One launch packs the operands of all four blocks from the current weights: 'raw' forward planes (identity
        BN, epilogue relu(acc + bias)) and the tap-flipped, channel-transposed dgrad planes of blocks 2-4."""
        p = self.p
        arr = C.c_void_p * 4
        self._pack_args = (arr(*[p[f"conv{i}_kernel"].data_ptr() for i in range(1, 5)]),
                           arr(*[p[f"conv{i}_bias"].data_ptr() for i in range(1, 5)]),
                           arr(*[t.data_ptr() for t in self.wraw]), arr(*[t.data_ptr() for t in self.eraw]),
                           arr(*[t.data_ptr() if t is not None else 0 for t in self.wdg]),
                           arr(*[t.data_ptr() if t is not None else 0 for t in self.edg]))
        k, bi, wr, er, wd, ed = self._pack_args
        plan.launch(self.lib.vm_pack_train, "vm_pack_train", k, bi, self.filters, wr, er, wd, ed, st)

    def _set_masks(self, nb, masks):
        """Fill the static SpatialDropout1D mask buffers for this step; returns whether dropout is active.  One
        Bernoulli draw per (clip, channel), scaled by 1 / keep (keras guards 0 < rate < 1: identity otherwise)."""
        if masks is not None:
            active = any(m is not None for m in masks)
            for buf, m in zip(self.maskbuf, masks):
                if active:
                    buf.copy_(m) if m is not None else buf.fill_(1.0)
            self.masks = list(self.maskbuf) if active else [None] * 4
            return active
        if not (0.0 < self.dropout < 1.0):
            self.masks = [None] * 4
            return False
        keep = 1.0 - self.dropout
        for buf in self.maskbuf:
            buf.copy_((torch.rand(buf.shape, generator=self.gen, device=self.device) < keep).to(torch.float32) / keep)
        self.masks = list(self.maskbuf)
        return True

    # ------------------------------------------------------------------ forward (train mode)
    def forward_train(self, x, groups, masks=None, update_moving=True, dense=True):
        """x: CUDA fp32 (NB, L).  Returns embeddings (NB, E).  Keeps everything backward needs.  dense=False stops after
        GlobalMaxPool1D (the siamese step's fused head call applies the embedding Dense itself)."""
        nb, length = x.shape
        self._buffers(nb, length, groups)
        if x.data_ptr() != self.xin.data_ptr():
            self.xin.copy_(x)
        self.x_in = self.xin
        self.groups = groups
        dropout_on = self._set_masks(nb, masks)
        key = ("fwd", dropout_on, bool(update_moving), self.sync_allreduce is not None, self.sync_peers is not None,
               bool(dense), torch.cuda.current_stream().cuda_stream)
        plan = self._plans.get(key)
        if plan is None:
            plan = self._plans[key] = self._build_forward_plan(nb, length, groups, update_moving, dense)
        plan.run()
        return self.embv

    def _build_forward_plan(self, nb, length, groups, update_moving, dense=True):
        lib, plan, st = self.lib, _Plan(), _stream()
        self._plan_pack(plan, st)
        c, ls = self.channels, self.ls
        eps, mom = C.c_float(BN_EPS), C.c_float(BN_MOMENTUM)
        for b in range(4):
            gamma, beta = _ptr(self.p[f"bn{b + 1}_gamma"]), _ptr(self.p[f"bn{b + 1}_beta"])
            if b == 0:
                plan.launch(lib.vm_conv1_train_fwd, "train conv block 1", _ptr(self.xin), nb, length, c[0], self.pools[0],
                            _ptr(self.wraw[0]), _ptr(self.eraw[0]), gamma, _ptr(self.U16[0]), _ptr(self.EXT[0]),
                            _ptr(self.stat[0]), self.precision, st)
            else:
                second = self.XQ[b - 1] if self.precision == 2 else self.XL[b - 1]
                plan.launch(lib.vm_conv3_train_fwd, f"train conv block {b + 1}", _ptr(self.X[b - 1]),
                            _ptr(second), nb, ls[b], c[b - 1], c[b], _ptr(self.wraw[b]), _ptr(self.eraw[b]),
                            gamma, _ptr(self.U16[b]), _ptr(self.EXT[b]), _ptr(self.stat[b]), self.precision, st)
            mm = self.moving[f"bn{b + 1}_mean"] if update_moving else None
            mv = self.moving[f"bn{b + 1}_var"] if update_moving else None
            if self.sync_allreduce is None:
                plan.launch(lib.vm_bn_stats_finalize, "vm_bn_stats_finalize", _ptr(self.stat[b]), self.stat_rows[b], nb,
                            groups, ls[b], c[b], gamma, beta, eps, mom, _ptr(mm), _ptr(mv), _ptr(self.bnc[b]),
                            _ptr(self.red), st)
            elif self.sync_peers is not None:
                # reduction, exchange over peer memory and constants in one launch (column finishers, vm_p2p.cuh)
                pe = self.sync_peers
                k = groups * c[b] * 2
                loc, glo = self.sums[0][:k], self.sums[1][:k]
                plan.host(pe.bump)
                count = float(self.sync_world) * (nb // groups) * ls[b]   # equal shards on every rank
                plan.launch(lib.vm_bn_stats_finalize_peers, "vm_bn_stats_finalize_peers", _ptr(self.stat[b]),
                            self.stat_rows[b], nb, groups, c[b], gamma, beta, eps, mom, _ptr(mm), _ptr(mv),
                            _ptr(self.bnc[b]), _ptr(self.red), pe.peers, pe.rank, pe.world, pe.seq, C.c_double(count),
                            _ptr(loc), _ptr(glo), st)
            else:
                sums = self.sums[1][:groups * c[b] * 2]
                plan.launch(lib.vm_bn_stats_sums, "vm_bn_stats_sums", _ptr(self.stat[b]), self.stat_rows[b], nb, groups,
                            c[b], _ptr(self.red), _ptr(sums), st)
                plan.host(lambda sums=sums: self.sync_allreduce(sums))
                count = float(self.sync_world) * (nb // groups) * ls[b]   # equal shards on every rank
                plan.launch(lib.vm_bn_stats_from_sums, "vm_bn_stats_from_sums", _ptr(sums), C.c_double(count), groups,
                            c[b], gamma, beta, eps, mom, _ptr(mm), _ptr(mv), _ptr(self.bnc[b]), st)
            if b < 3:
                plan.launch(lib.vm_bn_pool_fwd, "vm_bn_pool_fwd", _ptr(self.EXT[b]), nb, ls[b + 1], c[b], groups,
                            _ptr(self.bnc[b]), _ptr(self.masks[b]), _ptr(self.X[b]), _ptr(self.XL[b]), _ptr(self.XQ[b]),
                            st)
            else:
                plan.launch(lib.vm_bn_gmax_fwd, "vm_bn_gmax_fwd", _ptr(self.EXT[3]), nb, ls[4], c[3], groups,
                            _ptr(self.bnc[3]), _ptr(self.masks[3]), _ptr(self.gmax), _ptr(self.jstar), st)
        if dense:
            plan.launch(lib.vm_dense_fwd, "vm_dense_fwd", _ptr(self.gmax), nb, c[3], _ptr(self.p["dense_kernel"]),
                        _ptr(self.p["dense_bias"]), self.emb, _ptr(self.embv), st)
        return plan

    def bucket_bounds(self):
        """Element ranges of the flat gradient buffer in the order the backward pass completes them: the embedding
        Dense + head, then blocks 4, 3, 2, 1 (conv kernel, conv bias, BN gamma, BN beta are adjacent in Keras order)."""
        first = {i: self.layout[f"conv{i}_kernel"][0] for i in range(1, 5)}
        dense = self.layout["dense_kernel"][0]
        return [(dense, self.nparams), (first[4], dense), (first[3], first[4]), (first[2], first[3]), (0, first[2])]

    def set_gradient_buckets(self, enabled=True):
        """Data parallel: all-reduce the gradient per block, as soon as that block's weight gradient is written, on the
        process group's stream while the backward pass continues; ``siamese_step`` / ``classifier_step`` wait for the
        buckets before the optimizer step.  Replaces the single all-reduce after backward (``allreduce=`` of the steps)."""
        if enabled:
            from .parallel import GradientBuckets
            self.grad_buckets = GradientBuckets(self.grad, self.bucket_bounds())
        else:
            self.grad_buckets = None
        self._plans = {k: v for k, v in self._plans.items() if k[0] != "bwd"}

    def set_sync_bn(self, allreduce, world, peers=None):
        """allreduce(tensor): in-place SUM over ranks (torch.distributed.all_reduce); world: number of ranks, each
        feeding the same number of clips per step.  allreduce=None restores per-rank statistics.  ``peers``
        (parallel.PeerExchange) replaces the collective calls by the kernels that sum over NVLink peer memory."""
        self.sync_allreduce = allreduce
        self.sync_world = int(world) if allreduce is not None else 1
        self.sync_peers = peers if allreduce is not None else None
        self._plans = {}

    # ------------------------------------------------------------------ backward of the encoder
    def backward_encoder(self, d_emb, dense=True):
        """d_emb (NB, E) (already multiplied by the loss scale).  Fills self.g for all encoder parameters.  dense=False:
        the caller has already written self.d_gmax and the Dense gradients (the siamese step's fused head call)."""
        if dense and d_emb.data_ptr() != self.d_emb.data_ptr():
            self.d_emb.copy_(d_emb)
        key = ("bwd", self.masks[0] is not None, self.sync_allreduce is not None, self.sync_peers is not None,
               self.grad_buckets is not None, bool(dense), bool(self.overlap_wgrad),
               torch.cuda.current_stream().cuda_stream)
        plan = self._plans.get(key)
        if plan is None:
            plan = self._plans[key] = self._build_backward_plan(self.d_emb.shape[0], dense)
        plan.run()

    def _build_backward_plan(self, nb, dense=True):
        lib, plan, st = self.lib, _Plan(), _stream()
        c, ls, g, groups = self.channels, self.ls, self.g, self.groups
        if dense:
            plan.launch(lib.vm_dense_bwd, "vm_dense_bwd", _ptr(self.gmax), _ptr(self.d_emb), _ptr(self.p["dense_kernel"]),
                        nb, c[3], self.emb, _ptr(g["dense_kernel"]), _ptr(g["dense_bias"]), _ptr(self.d_gmax), st)
        buckets = self.grad_buckets
        if buckets is not None:    # Dense + head gradients are complete (the head's were written before this plan)
            plan.host(lambda: buckets.launch(0))
        bp = self.bwd_precision
        side = None
        if self.overlap_wgrad:
            if self._side is None:
                self._side = torch.cuda.Stream(device=self.device)
            side = self._side
            main = torch.cuda.current_stream()
            st2 = C.c_void_p(side.cuda_stream)
            du_ready = [torch.cuda.Event() for _ in range(4)]     # block b's dU written (main stream)
            wgrad_done = [torch.cuda.Event() for _ in range(4)]   # block b's weight gradient finished (side stream)
        for b in (3, 2, 1, 0):
            n_u = nb * ls[b] * c[b]
            du_hi = self.dU[b % 2][0][:n_u]
            du_lo = self.dU[b % 2][1][:n_u] if bp == 3 else None
            gabs = _ptr(self.gabs[b:b + 1])
            if side is not None and b <= 1:    # this block's dU set was last read by block b+2's weight gradient
                plan.host(lambda e=wgrad_done[b + 2]: main.wait_event(e))
            if b == 3:
                dy, dg, js = None, self.d_gmax, self.jstar
            else:
                dy, dg, js = self.dX, None, None
            grads = (_ptr(g[f"bn{b + 1}_gamma"]), _ptr(g[f"bn{b + 1}_beta"]))
            # blocks 1-3: the sums of dy and dy * xhat were taken by the dgrad epilogue of the block above
            pre_rows = lib.vm_conv3_train_rows_per_clip(ls[b + 1]) if b < 3 else 0
            part = self.redp[b] if b < 3 else self.scr2
            if self.sync_allreduce is None:
                plan.launch(lib.vm_bn_bwd, f"vm_bn_bwd block {b + 1}", _ptr(self.U16[b]), _ptr(self.EXT[b]), _ptr(dy),
                            _ptr(dg), _ptr(js), nb, ls[b], c[b], groups, self.pools[b], _ptr(self.bnc[b]),
                            _ptr(self.masks[b]), _ptr(part), _ptr(self.bwc[b]), *grads, gabs, _ptr(du_hi),
                            _ptr(du_lo), _ptr(self.scr1), _ptr(g[f"conv{b + 1}_bias"]), _ptr(self.red), pre_rows, st)
            elif self.sync_peers is not None:
                pe = self.sync_peers
                k = groups * c[b] * 2
                loc, glo = self.sums[0][:k], self.sums[1][:k]
                plan.host(pe.bump)
                count = float(self.sync_world) * (nb // groups) * ls[b]
                plan.launch(lib.vm_bn_bwd_peers, f"vm_bn_bwd_peers block {b + 1}", _ptr(self.U16[b]), _ptr(self.EXT[b]),
                            _ptr(dy), _ptr(dg), _ptr(js), nb, ls[b], c[b], groups, self.pools[b], _ptr(self.bnc[b]),
                            _ptr(self.masks[b]), _ptr(part), _ptr(self.bwc[b]), *grads, gabs, _ptr(du_hi), _ptr(du_lo),
                            _ptr(self.scr1), _ptr(g[f"conv{b + 1}_bias"]), _ptr(self.red), pre_rows, pe.peers, pe.rank,
                            pe.world, pe.seq, C.c_double(count), _ptr(loc), _ptr(glo), st)
            else:
                k = groups * c[b] * 2
                loc, glo = self.sums[0][:k], self.sums[1][:k]
                plan.launch(lib.vm_bn_bwd_sums, f"vm_bn_bwd_sums block {b + 1}", _ptr(self.EXT[b]), _ptr(dy), _ptr(dg),
                            _ptr(js), nb, ls[b], c[b], groups, self.pools[b], _ptr(self.bnc[b]), _ptr(self.masks[b]),
                            _ptr(part), gabs, _ptr(self.red), _ptr(loc), pre_rows, st)

                def share(loc=loc, glo=glo):
                    glo.copy_(loc)
                    self.sync_allreduce(glo)
                plan.host(share)
                count = float(self.sync_world) * (nb // groups) * ls[b]
                plan.launch(lib.vm_bn_bwd_from_sums, f"vm_bn_bwd_from_sums block {b + 1}", _ptr(loc), _ptr(glo),
                            C.c_double(count), _ptr(self.U16[b]), _ptr(dy), _ptr(dg), _ptr(js), nb, ls[b], c[b], groups,
                            self.pools[b], _ptr(self.bnc[b]), _ptr(self.masks[b]), _ptr(self.bwc[b]), *grads, gabs,
                            _ptr(du_hi), _ptr(du_lo), _ptr(self.scr1), _ptr(g[f"conv{b + 1}_bias"]), _ptr(self.red), st)
            if b > 0:
                # dgrad: dX_{b-1} = conv3(dU_b, flipped/transposed W_b), fp32 (NB, ls[b], c[b-1])
                # ... and, in its epilogue, the BatchNorm-backward sums of block b (dX is that block's pooled gradient)
                plan.launch(lib.vm_conv3_dgrad, f"dgrad block {b + 1}", _ptr(du_hi), _ptr(du_lo), nb, ls[b], c[b],
                            c[b - 1], _ptr(self.wdg[b]), _ptr(self.edg[b]), gabs, _ptr(self.dX), bp,
                            _ptr(self.EXT[b - 1]), _ptr(self.bnc[b - 1]), _ptr(self.masks[b - 1]), groups,
                            _ptr(self.redp[b - 1]), _ptr(self.gabs[b - 1:b]), st)
            stw = st
            if side is not None:
                # the weight gradient starts when dgrad has left the SMs (both kernels take a whole SM's shared memory):
                # it then runs beside the next block's BatchNorm / ReLU backward pass
                def fork(e=du_ready[b]):
                    e.record(main)
                    side.wait_event(e)
                plan.host(fork)
                stw = st2

            def share(i):   # block's gradients complete (its BatchNorm / bias gradients precede du_ready): all-reduce them
                if side is None:
                    buckets.launch(i)
                else:
                    with torch.cuda.stream(side):
                        buckets.launch(i)
            if b == 0:
                plan.launch(lib.vm_wgrad1, "vm_wgrad1", _ptr(self.xin), _ptr(du_hi), _ptr(du_lo), nb, ls[0], c[0], bp,
                            gabs, _ptr(self.wpart), self.wpart.numel() * 4, _ptr(g["conv1_kernel"]), stw)
                if buckets is not None:
                    plan.host(lambda: share(4))
            else:
                x_lo = self.XL[b - 1]
                wp = bp if x_lo is not None else 1      # forward precision 1 keeps no second activation plane
                plan.launch(lib.vm_wgrad3, f"vm_wgrad3 block {b + 1}", _ptr(self.X[b - 1]), _ptr(x_lo), _ptr(du_hi),
                            _ptr(du_lo if wp == 3 else None), nb, ls[b], c[b - 1], c[b], wp, gabs, _ptr(self.wpart),
                            self.wpart.numel() * 4, _ptr(g[f"conv{b + 1}_kernel"]), stw)
                if buckets is not None:    # block b+1's gradients are complete: share them while dgrad and the
                    plan.host(lambda i=4 - b: share(i))            # blocks below keep the device busy
            if side is not None:
                plan.host(lambda e=wgrad_done[b]: e.record(side))
        if side is not None:    # the optimizer (and the next step's forward pass) follow the last weight gradient
            plan.host(lambda e=wgrad_done[0]: main.wait_event(e))
        return plan

    # ------------------------------------------------------------------ optimizer
    def apply_gradients(self, world=1):
        opt = self.optimizer
        self.iterations += 1
        t = self.iterations
        lr = opt.lr
        if opt.decay > 0:
            lr = lr * (1.0 / (1.0 + opt.decay * (t - 1)))
        lr_t = lr * np.sqrt(1.0 - opt.beta_2 ** t) / (1.0 - opt.beta_1 ** t)
        rc = self.lib.vm_adam_step(_ptr(self.flat), _ptr(self.grad), _ptr(self.adam_m), _ptr(self.adam_v),
                                   self.nparams, _ptr(self.sumsq), C.c_float(1.0 / (self.loss_scale * world)),
                                   C.c_float(opt.clipnorm if opt.clipnorm else 0.0), C.c_float(lr_t),
                                   C.c_float(opt.beta_1), C.c_float(opt.beta_2), C.c_float(opt.epsilon), _stream())
        _check(rc, "vm_adam_step")
        opt.iterations = t
        self.eng._packed = False

    # ------------------------------------------------------------------ full steps
    def _stage(self, src, dst):
        """Host or device data -> the static device buffer ``dst`` without blocking the host: numpy batches (float64
        from the batcher, voicemap/librispeech.py:103) are cast into a ring of pinned buffers and copied asynchronously,
        so the host can issue the next step while this one runs."""
        if isinstance(src, torch.Tensor):
            t = src.reshape(dst.shape)
            dst.copy_(t, non_blocking=True)
            return
        a = np.asarray(src)
        if a.ndim == dst.dim() + 1:
            a = a[..., 0]                      # (N, L, 1) -> (N, L)
        a = a.reshape(dst.shape)
        k = self._pin_next
        self._pin_next = (k + 1) % len(self._pin)
        slot = self._pin[k]
        if slot is not None and slot[1] is not None:
            slot[1].synchronize()              # the copy that used this slot three stagings ago
        if slot is None or slot[0].numel() < a.size:
            slot = self._pin[k] = [torch.empty(a.size, dtype=torch.float32).pin_memory(), None]
        view = slot[0][:a.size].view(dst.shape)
        _host_copy(view.numpy(), a)            # cast + gather in one pass (row blocks on a few threads when large)
        dst.copy_(view, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        slot[1] = ev

    @staticmethod
    def _clip_shape(x):
        shape = tuple(x.shape)
        if len(shape) == 3:
            if shape[2] != 1:
                raise ValueError(f"clips must have one channel, got shape {shape}")
            shape = shape[:2]
        return shape

    def siamese_step(self, x1, x2, y, apply=True, masks=None, allreduce=None, world=1):
        """One train_on_batch of the siamese model.  Returns (loss, accuracy) as 0-d device tensors (no host sync)."""
        n, length = self._clip_shape(x1)
        if self._clip_shape(x2) != (n, length):
            raise ValueError("both inputs of a pair batch must have the same shape")
        self._buffers(2 * n, length, 2)
        self._stage(x1, self.xin[:n])
        self._stage(x2, self.xin[n:])
        self._stage(np.asarray(y, dtype=np.float32).reshape(-1) if not isinstance(y, torch.Tensor) else y, self.yin)
        self.forward_train(self.xin, groups=2, masks=masks, dense=False)
        key = ("head", torch.cuda.current_stream().cuda_stream)
        plan = self._plans.get(key)
        if plan is None:
            plan = self._plans[key] = self._build_siamese_head_plan(n)
        plan.run()
        self.backward_encoder(self.d_emb, dense=False)
        self._share_gradients(allreduce)
        if apply:
            self.apply_gradients(world)
        out = self.lossv.clone()        # one copy: the next step overwrites the buffer while callers still hold these
        return out[0], out[1]

    def _share_gradients(self, allreduce):
        if self.grad_buckets is not None:
            self.grad_buckets.wait()          # issued per block inside the backward plan
        elif allreduce is not None:
            allreduce(self.grad)

    def _build_siamese_head_plan(self, n):
        lib, plan, st = self.lib, _Plan(), _stream()
        metric = self.model.distance_metric
        metric_id = {"uniform_euclidean": 0, "weighted_l1": 1}[metric]
        loss_id = 1 if self.loss == "contrastive_loss" else 2
        hw, hb = self.p["head_kernel"].reshape(-1), self.p["head_bias"]
        g = self.g
        # embedding Dense, distance layer, Dense(1, sigmoid), loss, accuracy and their backward pass down to d_gmax: one
        # call, two launches (per-pair kernel + Dense weight gradient / batch reductions)
        plan.launch(lib.vm_siamese_head_train, "vm_siamese_head_train", _ptr(self.gmax), n, self.channels[3], self.emb,
                    _ptr(self.p["dense_kernel"]), _ptr(self.p["dense_bias"]), metric_id, _ptr(hw), _ptr(hb),
                    _ptr(self.yin), loss_id, C.c_float(self.loss_scale), _ptr(self.embv), _ptr(self.prob),
                    _ptr(self.d_emb), _ptr(self.d_gmax), _ptr(self.pairrec), _ptr(g["dense_kernel"]),
                    _ptr(g["dense_bias"]), _ptr(g["head_kernel"]), _ptr(g["head_bias"]), _ptr(self.lossv), st)
        return plan

    def classifier_step(self, x, y_onehot, apply=True, masks=None, allreduce=None, world=1):
        """Encoder + Dense(softmax) + categorical cross-entropy.  The softmax head is adjacent to the hot path
         and runs as torch device ops."""
        n, length = self._clip_shape(x)
        self._buffers(n, length, 1)
        self._stage(x, self.xin)
        yt = torch.as_tensor(np.asarray(y_onehot, dtype=np.float32)).to(self.device, non_blocking=True) \
            if not isinstance(y_onehot, torch.Tensor) else y_onehot.to(self.device, dtype=torch.float32)
        emb = self.forward_train(self.xin, groups=1, masks=masks)
        hk, hb = self.p["head_kernel"], self.p["head_bias"]
        logits = emb @ hk + hb
        logp = torch.log_softmax(logits, dim=-1)
        lossv = -(yt * logp).sum(dim=-1).mean()
        dlogits = (torch.softmax(logits, dim=-1) - yt) * (self.loss_scale / n)
        self.g["head_kernel"].copy_(emb.t() @ dlogits)
        self.g["head_bias"].copy_(dlogits.sum(dim=0))
        self.d_emb.copy_(dlogits @ hk.t())
        self.backward_encoder(self.d_emb)
        self._share_gradients(allreduce)
        if apply:
            self.apply_gradients(world)
        acc = (logits.argmax(dim=-1) == yt.argmax(dim=-1)).to(torch.float32).mean()
        return lossv, acc

    # ------------------------------------------------------------------ weights back to the host model
    def sync_to_model(self):
        enc = self.encoder_model
        for name in enc.weights:
            if name in self.p and name not in self.head_names:
                enc.weights[name] = self.p[name].detach().cpu().numpy().copy()
            elif name in self.moving:
                enc.weights[name] = self.moving[name].detach().cpu().numpy().copy()
        target = self.model.head_weights if self.kind == "siamese" else enc.weights
        for name in self.head_names:
            target[name] = self.p[name].detach().cpu().numpy().copy()
        if self.kind == "siamese":
            self.model._head_dev = None

    def time_siamese_step(self, x1, x2, y, steps=10):
        """Live per-launch durations of a siamese step (CUDA events between the C-ABI calls, averaged over ``steps``
        steps after one warm-up step): [(what, ms), ...] in launch order, Adam last."""
        overlap, self.overlap_wgrad = self.overlap_wgrad, False   # one stream: an event after a launch brackets it
        try:
            self.siamese_step(x1, x2, y)
            acc = None
            for _ in range(steps):
                timings = []
                run, _Plan.run = _Plan.run, lambda plan: plan.run_timed(timings)
                try:
                    self.siamese_step(x1, x2, y, apply=False)
                finally:
                    _Plan.run = run
                t0 = torch.cuda.Event(enable_timing=True)
                t0.record()
                self.apply_gradients()
                t1 = torch.cuda.Event(enable_timing=True)
                t1.record()
                timings.append(("vm_adam_step", t0, t1))
                torch.cuda.synchronize()
                ms = [(w, a.elapsed_time(b)) for w, a, b in timings]
                acc = ms if acc is None else [(w, t + u) for (w, t), (_, u) in zip(acc, ms)]
        finally:
            self.overlap_wgrad = overlap
        return [(w, t / steps) for w, t in acc]

    # ------------------------------------------------------------------ views for tests / diagnostics
    def relu_pattern(self, b):
        """(NB, L_b, C_b) bool: where block b+1's un-pooled activation of the last step was positive."""
        return (self.U16[b] & 0x7FFF) != 0

    def activation(self, b):
        """(NB, L_b, C_b) fp32: block b+1's un-pooled u = relu(conv + bias) as kept for the backward pass (fp16)."""
        return (self.U16[b] & 0x7FFF).view(torch.float16).to(torch.float32)

    def argmax_flags(self, b):
        return self.U16[b] < 0          # bit 15 of the int16 word

    def block_gradient(self, b):
        """(NB, L_b, C_b) fp32: dLoss/d(conv output) of block b+1 as the backward kernels consumed it (planes summed,
        gradient scale and loss scale taken out).  Only the block written last is still in the buffer (block 1 after
        a full backward)."""
        n_u = self.U16[b].numel()
        planes = self.dU[b % 2][:, :n_u].to(torch.float32).sum(dim=0)
        absmax = self.gabs[b:b + 1].view(torch.float32).item()
        if absmax > 0:
            import math
            planes = planes / math.ldexp(1.0, 6 - math.frexp(absmax)[1])
        return (planes / self.loss_scale).view(self.U16[b].shape)

    def gradients(self):
        """Unscaled gradients as numpy arrays (tests)."""
        return OrderedDict((k, (v / self.loss_scale).detach().cpu().numpy()) for k, v in self.g.items())


# ----------------------------------------------------------------------------------------------------------------
# fit_generator (Keras 2.2.2 semantics for the subset the scripts use)
# ----------------------------------------------------------------------------------------------------------------
_COPY_POOL = None
_COPY_THREADS = None      # decided by measurement at the first large copy: 1 = plain assignment, else pool width


def _host_copy(dst, src):
    """dst[...] = src for numpy arrays of equal shape (any source dtype).  One thread moves ~6-11 GB/s, which makes the
    staging of a 128-pair float32 batch (12 MB) a full millisecond of every end-to-end step; numpy releases the GIL while
    it copies, so large batches may go in row blocks on up to four threads -- whether that is faster depends on the host
    (it is not on a busy 8-vCPU container), so the first large copy times both ways once and the faster one is kept."""
    global _COPY_POOL, _COPY_THREADS
    rows = dst.shape[0] if dst.ndim > 1 else 0
    if dst.size < (1 << 19) or rows < 8 or _COPY_THREADS == 1:
        dst[...] = src
        return

    def pooled():
        k = _COPY_POOL._max_workers
        step = -(-rows // k)

        def block(lo):
            dst[lo:lo + step] = src[lo:lo + step]
        list(_COPY_POOL.map(block, range(0, rows, step)))

    if _COPY_THREADS is None:
        import time
        from concurrent.futures import ThreadPoolExecutor
        try:
            cpus = len(os.sched_getaffinity(0))
        except AttributeError:
            cpus = os.cpu_count() or 1
        if cpus < 4:
            _COPY_THREADS = 1
            dst[...] = src
            return
        _COPY_POOL = ThreadPoolExecutor(max_workers=4)
        pooled()                                   # warm the threads and the pages
        t0 = time.perf_counter()
        dst[...] = src
        t1 = time.perf_counter()
        pooled()
        t2 = time.perf_counter()
        _COPY_THREADS = 4 if (t2 - t1) < 0.8 * (t1 - t0) else 1
        if _COPY_THREADS == 1:
            _COPY_POOL.shutdown(wait=False)
            _COPY_POOL = None
        return
    pooled()


def _next_batch(gen_iter, generator, step):
    if hasattr(generator, "__getitem__") and hasattr(generator, "__len__"):
        return generator[step % len(generator)]
    return next(gen_iter)


class _Prefetcher:
    """Pulls batches from the user's generator / Sequence on one background thread so that host-side batch
    construction (FLAC decode, pair sampling, whitening) overlaps the device step -- the role of Keras'
    ``workers`` / ``use_multiprocessing`` queue (experiments/train_siamese.py:71-72).  Order is preserved and the
    generator is only ever touched from that one thread."""

    def __init__(self, generator, depth=2):
        import queue
        import threading
        self.generator = generator
        self.is_seq = hasattr(generator, "__getitem__") and hasattr(generator, "__len__")
        self.it = None if self.is_seq else iter(generator)
        self.q = queue.Queue(maxsize=depth)
        self.step = 0
        self.stop = False
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        try:
            while not self.stop:
                item = _next_batch(self.it, self.generator, self.step)
                self.step += 1
                while not self.stop:
                    try:
                        self.q.put(item, timeout=0.1)
                        break
                    except Exception:
                        continue
        except BaseException as exc:  # surfaced to the training loop
            self.q.put(exc)

    def next(self):
        item = self.q.get()
        if isinstance(item, BaseException):
            if isinstance(item, StopIteration):
                raise StopIteration
            raise item
        return item

    def close(self):
        self.stop = True


def sync_bn_peers(trainer=None):
    """The peer-memory exchange for synchronised BatchNorm (parallel.PeerExchange), created once per process --
    or None when VOICEMAP_SYNCBN=nccl asks for the collective-call form (jobs beyond one node / 8 ranks)."""
    if os.environ.get("VOICEMAP_SYNCBN", "p2p").lower() == "nccl":
        return None
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_backend() != "nccl":
        return None                          # gloo (CPU tests of the host logic): collective calls
    global _PEERS
    if _PEERS is None:
        from .parallel import PeerExchange
        _PEERS = PeerExchange()
    return _PEERS


_PEERS = None


def _dist_info():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist, dist.get_world_size()
    return None, 1


def _producer_processes(workers, use_multiprocessing):
    """How many batch-producer processes to fork: Keras' ``workers`` when ``use_multiprocessing`` is set (the reference
    passes ``multiprocessing.cpu_count()``), capped by the cores this process may use; 0 = stay on the background
    thread."""
    import multiprocessing as mp
    if not use_multiprocessing or not workers or int(workers) <= 1 or 'fork' not in mp.get_all_start_methods():
        return 0
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    return max(0, min(int(workers), cores, 16)) if cores > 1 else 0


def fit_generator(model, generator, steps_per_epoch=None, epochs=1, verbose=1, callbacks=None, validation_data=None,
                  validation_steps=None, initial_epoch=0, workers=1, use_multiprocessing=False):
    """model.fit_generator(...) of experiments/train_siamese.py:65-94 / train_classifier.py:120-150.  ``workers`` > 1
    with ``use_multiprocessing=True`` (what both scripts pass) forks that many batch producers (``prefetch.py``)."""
    if model.loss is None or model.optimizer is None:
        raise RuntimeError("You must compile a model before training/testing. Use `model.compile(optimizer, loss)`.")
    is_seq = hasattr(generator, "__getitem__") and hasattr(generator, "__len__")
    if steps_per_epoch is None:
        if not is_seq:
            raise ValueError("`steps_per_epoch=None` is only valid for a generator based on the `Sequence` class.")
        steps_per_epoch = len(generator)
    trainer = getattr(model, "_trainer", None)
    if _dist_info()[1] > 1 and trainer is None:
        from .parallel import broadcast_weights_
        broadcast_weights_(model)            # ranks were initialised independently: start from rank 0's weights
    if (trainer is None or trainer.optimizer is not model.optimizer or trainer.loss != model.loss
            or trainer.bwd_precision != getattr(model, "train_bwd_precision", 1)):
        trainer = TrainEngine(model, model.optimizer, model.loss)
        model._trainer = trainer
    dist, world = _dist_info()
    allreduce = (lambda t: dist.all_reduce(t)) if world > 1 else None
    # BatchNorm sees the whole batch in the reference (one device); data-parallel ranks therefore share their batch
    # statistics unless the model opts out with ``model.sync_batchnorm = False``
    want_sync = world > 1 and getattr(model, "sync_batchnorm", True)
    trainer.set_sync_bn(allreduce if want_sync else None, world, peers=sync_bn_peers(trainer) if want_sync else None)
    trainer.set_gradient_buckets(world > 1)     # per-block asynchronous gradient all-reduce inside the backward pass
    callbacks = list(callbacks or [])
    for cb in callbacks:
        cb.set_model(model)
        cb.set_params(dict(epochs=epochs, steps=steps_per_epoch, verbose=verbose))
        cb.on_train_begin()
    producers = _producer_processes(workers, use_multiprocessing)
    if is_seq:
        prefetch = None                    # Sequences are indexed (and reshuffled) per epoch
    elif producers:
        from .prefetch import ProcessPrefetcher
        prefetch = ProcessPrefetcher(generator, producers)
    else:
        prefetch = _Prefetcher(generator)
    val_iter = None
    if validation_data is not None and not isinstance(validation_data, (tuple, list)):
        val_iter = iter(validation_data)
    history = []
    try:
        for epoch in range(initial_epoch, epochs):
            for cb in callbacks:
                cb.on_epoch_begin(epoch)
            losses, accs = [], []
            ordered = None
            if is_seq and producers and steps_per_epoch > 1:
                from .prefetch import SequencePrefetcher
                ordered = SequencePrefetcher(generator, [s % len(generator) for s in range(steps_per_epoch)], producers)
            try:
                for step in range(steps_per_epoch):
                    if ordered is not None:
                        batch = ordered.next()
                    else:
                        batch = prefetch.next() if prefetch is not None else generator[step % len(generator)]
                    x, y = batch[0], batch[1]
                    if trainer.kind == "siamese":
                        lv, acc = trainer.siamese_step(x[0], x[1], y, allreduce=allreduce, world=world)
                    else:
                        lv, acc = trainer.classifier_step(x, y, allreduce=allreduce, world=world)
                    losses.append(lv)
                    accs.append(acc)
            finally:
                if ordered is not None:
                    ordered.close()
            # data parallel: every rank reports the mean over the GLOBAL batches, so that callbacks steered by the logs
            # (ReduceLROnPlateau, ModelCheckpoint, EarlyStopping) take the same decisions on every rank
            from .parallel import global_mean
            logs = {"loss": global_mean(float(torch.stack(losses).sum().item()), len(losses))}
            if "accuracy" in (model.metrics or []) or "acc" in (model.metrics or []):
                logs["acc"] = global_mean(float(torch.stack(accs).sum().item()), len(accs))
            trainer.sync_to_model()
            if validation_data is not None:
                vl, va = _validate(model, trainer, validation_data, val_iter, validation_steps)
                logs["val_loss"] = vl
                if "acc" in logs:
                    logs["val_acc"] = va
            if is_seq:
                generator.on_epoch_end()
            for cb in callbacks:
                cb.on_epoch_end(epoch, logs)
            if verbose:
                print(f"Epoch {epoch + 1}/{epochs} - " + " - ".join(f"{k}: {v:.4f}" for k, v in logs.items()))
            history.append(dict(logs))
    finally:
        if prefetch is not None:    # producers (thread or processes) stop even when a step raised
            prefetch.close()
    for cb in callbacks:
        cb.on_train_end()
    return history


def _validate(model, trainer, validation_data, val_iter, validation_steps):
    """Eval-mode loss / accuracy (moving BN statistics, no dropout) over the validation batches, Keras-style
    sample-weighted means.  Siamese: ``model.test_on_batch`` (encoder + fused head/loss kernel on the device).
    Classifier: softmax probabilities from the device, the cross-entropy of the (N, classes) matrix on the host (the
    classifier head is adjacent to the path)."""
    if val_iter is None:
        batches = [validation_data]
    else:
        if validation_steps is None:
            raise ValueError("`validation_steps` is required when validation_data is a generator")
        batches = (next(val_iter) for _ in range(validation_steps))
    tot, tl, ta = 0, 0.0, 0.0
    for batch in batches:
        x, y = batch[0], np.asarray(batch[1], dtype=np.float32)
        if trainer.kind == "siamese":
            # encoder, head and loss on the device: one 2N-clip launch + the fused head/loss kernel of the train step
            metrics, model.metrics = model.metrics, ["accuracy"]
            try:
                lv, acc = model.test_on_batch(x, y)
            finally:
                model.metrics = metrics
            n = y.reshape(-1).shape[0]
        else:
            pred = model.predict(x).astype(np.float32)
            n = pred.shape[0]
            pc = np.clip(pred / pred.sum(axis=-1, keepdims=True), 1e-7, 1 - 1e-7)
            lv = float(np.mean(-(y * np.log(pc)).sum(axis=-1)))
            acc = float(np.mean(pred.argmax(axis=-1) == y.argmax(axis=-1)))
        tot += n
        tl += lv * n
        ta += acc * n
    if _dist_info()[1] > 1:
        from .parallel import global_mean
        return global_mean(tl, tot), global_mean(ta, tot)
    return tl / max(tot, 1), ta / max(tot, 1)
