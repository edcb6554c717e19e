"""Multi-GPU plumbing (one process per GPU, ``torch.distributed``): the hot path shards by independent clips, so
inference needs no data-path collective; training all-reduces ONE flat fp32 gradient buffer per step
(SURVEY.md 8(e)).  Everything here is backend-agnostic host logic (NCCL on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def world():
    """(rank, world_size) -- (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def init_from_env(backend="nccl", device=None):
    """Initialise the default process group from RANK / WORLD_SIZE / MASTER_* (torchrun); no-op for one process."""
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    if ws <= 1 or dist.is_initialized():
        return world()
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    kwargs = {}
    if backend == "nccl" and device is not None:
        kwargs["device_id"] = device
    dist.init_process_group(backend, **kwargs)
    return world()


def shard_bounds(n_items, rank=None, world_size=None):
    """Contiguous shard [lo, hi) of ``n_items`` independent units for this rank; sizes differ by at most one and
    both members of a verification pair stay on one rank when the caller shards PAIRS."""
    if rank is None or world_size is None:
        rank, world_size = world()
    base, rem = divmod(n_items, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_sum_(flat):
    """In-place sum of the flat gradient buffer over all ranks (a single bucket: the 4.1 MB message is
    latency-bound on NVLink)."""
    if world()[1] > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    return flat


class GradientBuckets:
    """The flat gradient buffer cut into contiguous buckets, each all-reduced (sum) as soon as the backward pass has
    finished writing it, while the rest of the backward pass keeps running (one flat buffer,
    overlapped with the tail of backward).  ``bounds`` = [(lo, hi), ...] element ranges in the order in which they become
    complete (the encoder's blocks finish last-layer-first).  ``launch(i)`` issues bucket i asynchronously: with NCCL
    the collective runs on the process group's own stream, ordered after everything already enqueued on the current
    stream; ``wait()`` orders the current stream after all outstanding buckets (no host block with NCCL) -- call it
    before the optimizer reads the buffer.  One process: both are no-ops."""

    def __init__(self, flat, bounds):
        self.flat = flat
        self.bounds = [(int(lo), int(hi)) for lo, hi in bounds]
        covered = sorted(self.bounds)
        if covered[0][0] != 0 or covered[-1][1] != flat.numel() or any(a[1] != b[0] for a, b in zip(covered, covered[1:])):
            raise ValueError("gradient buckets must tile the flat buffer exactly once")
        self.views = [flat[lo:hi] for lo, hi in self.bounds]
        self.works = []
        self.bytes_per_step = flat.numel() * flat.element_size()

    def launch(self, i):
        if world()[1] > 1:
            self.works.append(dist.all_reduce(self.views[i], op=dist.ReduceOp.SUM, async_op=True))

    def wait(self):
        for work in self.works:
            work.wait()
        self.works = []


class PeerExchange:
    """Peer-memory exchange buffers for the cross-rank BatchNorm sums (``vm_bn_stats_finalize_peers`` / ``vm_bn_bwd_peers``,
    csrc/vm_p2p.cuh): every rank of the node allocates one buffer in libvoicemap_b200.so, publishes its CUDA IPC handle
    through ``torch.distributed`` (64 bytes per rank, once) and maps the others'.  The sums then travel over NVLink
    inside the consuming kernel -- no collective call, no host involvement per step.  ``seq`` is the call counter every
    rank advances in lockstep (``bump``) before each exchange; its ctypes object is handed to the launches, so recorded
    launch plans see the current value.  At most 8 ranks, one node."""

    def __init__(self):
        import ctypes as C
        from . import _lib
        self.lib = _lib.load()
        rank, ws = world()
        if ws > 8:
            raise _lib.VoicemapB200Error("peer-memory BatchNorm exchange supports at most 8 ranks of one node; set "
                                         "VOICEMAP_SYNCBN=nccl for larger jobs")
        own = C.c_void_p()
        _lib.check(self.lib.vm_p2p_alloc(C.byref(own)), "vm_p2p_alloc")
        handle = C.create_string_buffer(64)
        _lib.check(self.lib.vm_p2p_export(own, handle), "vm_p2p_export")
        handles = [None] * ws
        dist.all_gather_object(handles, handle.raw)
        self.own, self.imported, ptrs = own, [], []
        for r, h in enumerate(handles):
            if r == rank:
                ptrs.append(own.value)
            else:
                q = C.c_void_p()
                _lib.check(self.lib.vm_p2p_import(h, C.byref(q)), f"vm_p2p_import (rank {r})")
                self.imported.append(q)
                ptrs.append(q.value)
        self.peers = (C.c_void_p * ws)(*ptrs)
        self.rank, self.world = rank, ws
        self.seq = C.c_uint32(0)
        dist.barrier()                       # every buffer is zeroed and mapped before the first exchange

    def bump(self):
        self.seq.value += 1

    def close(self):
        for q in self.imported:
            self.lib.vm_p2p_unimport(q)
        self.imported = []
        if self.own is not None:
            self.lib.vm_p2p_free(self.own)
            self.own = None


def broadcast_weights_(model, src=0):
    """Every rank takes rank ``src``'s weights (``model.get_weights()`` / ``set_weights``).  The reference's builders
    have no seed argument (voicemap/models.py:6,44), so independently launched ranks would start data-parallel
    training from different random initialisations and -- sharing gradients but not weights -- never agree.  One flat
    float32 buffer, one broadcast; a no-op for one process.  With NCCL the buffer travels through the current CUDA
    device, with gloo through host memory."""
    if world()[1] <= 1:
        return model
    import numpy as np
    weights = model.get_weights()
    flat = torch.from_numpy(np.concatenate([np.asarray(w, dtype=np.float32).reshape(-1) for w in weights]))
    if dist.get_backend() == "nccl":
        flat = flat.cuda()
    dist.broadcast(flat, src=src)
    flat = flat.cpu().numpy()
    out, at = [], 0
    for w in weights:
        out.append(flat[at:at + w.size].reshape(w.shape).copy())
        at += w.size
    model.set_weights(out)
    return model


def _scalar_device(device):
    """Where small host-side scalars travel: NCCL only moves CUDA tensors, gloo host ones."""
    if device is not None:
        return device
    if dist.is_available() and dist.is_initialized() and dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return "cpu"


def max_over_ranks(value, device=None):
    """Timing reduction of the bench: the slowest rank defines the step time."""
    device = _scalar_device(device)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if world()[1] > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def global_mean(local_sum, local_count, device=None):
    """Mean of a per-sample quantity over the GLOBAL batch (the loss mean must not be a mean of rank means when
    shards are uneven)."""
    device = _scalar_device(device)
    t = torch.tensor([float(local_sum), float(local_count)], dtype=torch.float64, device=device)
    if world()[1] > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t[0].item() / max(t[1].item(), 1.0))


def gather_rows(x):
    """All-gather of per-rank embedding rows (k-way evaluation across ranks); rows may differ per rank."""
    rank, ws = world()
    if ws == 1:
        return x
    counts = [torch.zeros(1, dtype=torch.int64, device=x.device) for _ in range(ws)]
    dist.all_gather(counts, torch.tensor([x.shape[0]], dtype=torch.int64, device=x.device))
    counts = [int(c.item()) for c in counts]
    pad = max(counts)
    buf = torch.zeros((pad,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    buf[:x.shape[0]] = x
    out = [torch.empty_like(buf) for _ in range(ws)]
    dist.all_gather(out, buf)
    return torch.cat([o[:c] for o, c in zip(out, counts)], dim=0)
