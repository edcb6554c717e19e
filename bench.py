#!/usr/bin/env python
"""bench.py -- headline benchmark of the voicemap hot path on B200 (contract: see the task brief / DESIGN.md).

Metric (BASELINE.json): audio-seconds/sec embedded, 3 s clips, encoder forward (eval mode) of the baseline
1D-CNN (filters=128, embedding 64).  Workload = BASELINE config[1]: batch 256 clips per GPU, each clip entering
the network as 12000 samples (3 s @ 16 kHz after the reference's x4 decimation, voicemap/utils.py:29;
--length 48000 benches raw 16 kHz input instead).  Weak scaling: every rank embeds its own 256 clips, no
data-path collective.

    python bench.py [--gpus N --steps K --warmup W]         # this repo's CUDA path
    python bench.py --impl reference [...]                  # the reference's CPU path (oracle port), rank 0 only

One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FILTERS, EMB = 128, 64
CLIP_SECONDS = 3.0
METRIC = "audio_seconds_per_sec_embedded_3s_clips"
UNIT = "audio-s/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="clips per GPU per step")
    ap.add_argument("--length", type=int, default=12000, help="samples per clip entering the encoder")
    ap.add_argument("--precision", type=int, default=2, choices=[1, 2, 3],
                    help="2: fp16 product + fp8 (e5m2 pairs) correction product (parity mode, default); 3: fp16x3 split "
                         "(parity mode, tighter); 1: fp16x1 (throughput mode, not parity)")
    ap.add_argument("--sustained-seconds", type=float, default=2.0)
    ap.add_argument("--train-pairs", type=int, default=128,
                    help="GLOBAL pairs per training step (BASELINE config[2]: 128), sharded over the ranks")
    ap.add_argument("--train-steps", type=int, default=20)
    ap.add_argument("--no-train", action="store_true", help="skip the training-step record")
    ap.add_argument("--cpu-baseline-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return dict(hbm_gbs=float(d["hbm_gbs"]), bf16_tflops=float(d["bf16_tflops"]),
                    bf16_tflops_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0,
                source="fallback (B200_PROFILING.md)")


# ------------------------------------------------------------------------------------------------------------
# algorithmic work per clip: FLOP = 2*L*K*Cin*Cout, bytes = 4*(L*Cin + floor(L/p)*Cout)
# ------------------------------------------------------------------------------------------------------------
def block_work(length, f=FILTERS):
    blocks = []
    l, cin = length, 1
    for k, mult, pool in ((32, 1, 4), (3, 2, 2), (3, 3, 2), (3, 4, 2)):
        cout = mult * f
        lout = l // pool
        out_elems = lout * cout if len(blocks) < 3 else cout  # block 4 is fused with the global max
        blocks.append(dict(flop=2.0 * l * k * cin * cout, bytes=4.0 * (l * cin + out_elems)))
        l, cin = lout, cout
    return blocks


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []      # (arrival time, csv line)
        self.proc = None
        self.index = index
        self.window = None  # (t0, t1) perf_counter bounds of the timed region

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "10"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, bufsize=1)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
            t0 = time.perf_counter()
            while not self.rows and time.perf_counter() - t0 < 3.0:  # wait for the first sample (tool start-up)
                time.sleep(0.01)
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = self.rows
        if self.window is not None:
            inside = [r for r in rows if self.window[0] <= r[0] <= self.window[1]]
            # a very short timed region can fall between two samples: then take the nearest ones around it
            rows = inside if inside else sorted(rows, key=lambda r: abs(r[0] - self.window[1]))[:2]
        for _, r in rows:
            parts = [p.strip() for p in r.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


# ------------------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU path (oracle port: torch-CPU fp32 restatement of voicemap/models.py:6-41)
# ------------------------------------------------------------------------------------------------------------
def usable_cpus():
    """Host threads this process can really run: the affinity mask, capped by the cgroup CPU quota (a container limited
    to 8 CPUs on a 128-core host still reports 128 from os.cpu_count(), and 128 OpenMP threads on 8 CPUs thrash)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            with open(path) as f:
                fields = f.read().split()
            if path.endswith("cpu.max"):
                quota, period = fields[0], float(fields[1])
            else:
                quota = fields[0]
                with open("/sys/fs/cgroup/cpu/cpu.cfs_period_us") as f:
                    period = float(f.read().split()[0])
            if quota not in ("max", "-1"):
                n = max(1, min(n, int(np.ceil(float(quota) / period))))
            break
        except (OSError, ValueError, IndexError):
            continue
    return n


def time_oracle(length, clips, budget_s, steps=None, warmup=1):
    import torch
    from oracle import voicemap_oracle as O
    torch.set_num_threads(usable_cpus())
    params = O.init_encoder_params(FILTERS, EMB, seed=0, randomize_bn=True, random_bias=True)
    x = O.synthetic_clips(clips, length, seed=1234)
    for _ in range(warmup):
        O.encoder_forward(x, params, torch.float32)
    times = []
    t_start = time.perf_counter()
    while True:
        t0 = time.perf_counter()
        O.encoder_forward(x, params, torch.float32)
        times.append(time.perf_counter() - t0)
        if steps is not None and (len(times) >= steps or time.perf_counter() - t_start > 150.0):
            break  # bounded: at most `steps` passes and ~150 s of CPU time
        if steps is None and (time.perf_counter() - t_start > budget_s and len(times) >= 3):
            break
    med = float(np.median(times))
    return dict(value=clips * CLIP_SECONDS / med, unit=UNIT, cores=torch.get_num_threads(), kind="port",
                sample=f"{clips} clips x {length} samples per pass, {len(times)} timed passes (median), "
                       f"torch-CPU fp32 oracle, {warmup} warm-up"), med


def run_reference(args):
    """The reference's CPU path on the box's host cores (oracle port; the reference itself is Python 2.7 on Keras / TF,
    DESIGN.md section 7).  Step = one forward of the SAME batch as the GPU arm (--batch clips, BASELINE config[1]); the
    batch-8 case of BASELINE config[0] and the batch-64 case BASELINE.md section 2 asks for are timed beside it
    (3 passes each) and reported in ``other_batches``."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    clips = args.batch
    cb, med = time_oracle(args.length, clips, budget_s=0, steps=max(args.steps, 1), warmup=max(min(args.warmup, 2), 1))
    others = {}
    for n in (8, 64):
        if n != clips:
            ob, omed = time_oracle(args.length, n, budget_s=0, steps=3, warmup=1)
            others[f"batch{n}"] = dict(value=ob["value"], unit=UNIT, ms_per_pass=omed * 1e3, sample=ob["sample"])
    line = dict(metric=METRIC, value=cb["value"], unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=med * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f32", data="synthetic", impl="reference",
                config=dict(workload=f"baseline 1D-CNN encoder fwd (eval), filters={FILTERS}, emb={EMB}, batch={clips} "
                                     f"clips/step x {args.length} samples, CPU oracle port of voicemap/models.py:6-41 "
                                     f"(torch-CPU fp32, all host threads)"),
                cpu_baseline=cb, other_batches=others, host_cores=cb["cores"],
                e2e=dict(value=cb["value"], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------
# training step (BASELINE config[2]): siamese encoder + contrastive loss, 128 pairs per step, data parallel
# ------------------------------------------------------------------------------------------------------------
def bench_train(args, dev, rank, world, peaks):
    """One record per backward arithmetic: forward (train-mode BatchNorm over the GLOBAL batch), backward, per-block
    gradient all-reduce overlapped with the rest of backward, Keras-Adam(clipnorm=1).  STRONG scaling: the global batch
    of ``--train-pairs`` pairs (experiments/siamese_contrastive_loss.py:70,76-83 trains 64-pair batches; BASELINE
    config[2] names 128) is split over the ranks, so N ranks x 128/N pairs compute the single-device step.  Device timed
    (CUDA events around ``--train-steps`` steps, max over ranks); ``e2e`` feeds float32 numpy batches from host memory
    through the pinned staging ring every step, as ``fit_generator`` does."""
    import torch
    import torch.distributed as dist
    from voicemap_b200 import parallel
    from voicemap_b200.keras_compat import Adam
    from voicemap_b200.models import build_siamese_net, get_baseline_convolutional_encoder
    from voicemap_b200.training import TrainEngine
    from voicemap_b200.utils import contrastive_loss

    length, steps = args.length, max(args.train_steps, 1)
    if args.train_pairs % world != 0:
        return dict(skipped=f"{args.train_pairs} pairs do not split evenly over {world} ranks")
    pairs = args.train_pairs // world
    g = torch.Generator().manual_seed(4321 + rank)
    x1h = (0.038021 * torch.randn(pairs, length, generator=g)).numpy()
    x2h = (0.038021 * torch.randn(pairs, length, generator=g)).numpy()
    x1, x2 = torch.from_numpy(x1h).to(dev), torch.from_numpy(x2h).to(dev)
    y = np.concatenate([np.zeros(pairs // 2), np.ones(pairs - pairs // 2)]).astype(np.float32)
    yd = torch.from_numpy(y).to(dev)
    allreduce = parallel.allreduce_sum_ if world > 1 else None
    flops_fwd = sum(b["flop"] for b in block_work(length)) * 2 * args.train_pairs      # whole job, all ranks
    out = dict(metric="siamese_train_pairs_per_sec", unit="pairs/s", scaling="strong",
               config=dict(workload=f"siamese encoder + contrastive loss train step (fwd, bwd, all-reduce, Adam), "
                                    f"{args.train_pairs} pairs/step global = {pairs} pairs/GPU x {length} samples, "
                                    f"filters={FILTERS}, emb={EMB}, train-mode BatchNorm synchronised over the ranks",
                           parallelism=f"dp{world}: per-block gradient all-reduce (5 buckets, "
                                       f"{(1023808 + 2) * 4 / 1e6:.1f} MB fp32 per step, NCCL) overlapped with backward + 8 "
                                       f"BatchNorm statistics sums (<= 16 KB each) "
                                       f"{'inside the kernels over NVLink peer memory' if os.environ.get('VOICEMAP_SYNCBN', 'p2p').lower() != 'nccl' else 'as NCCL all-reduces'}"
                           if world > 1 else "dp1",
                           streams="weight gradients on a side stream beside the next block's BatchNorm/ReLU backward pass"),
               modes={})
    for name, bwd in (("bwd1", 1), ("bwd3", 3)):
        enc = get_baseline_convolutional_encoder(FILTERS, EMB, dropout=0.0)
        sia = build_siamese_net(enc, (length, 1))
        parallel.broadcast_weights_(sia)
        opt = Adam(clipnorm=1.0)
        sia.compile(loss=contrastive_loss, optimizer=opt)
        tr = TrainEngine(sia, opt, sia.loss, precision=3, bwd_precision=bwd)
        from voicemap_b200.training import sync_bn_peers
        peers = sync_bn_peers() if world > 1 else None
        tr.set_sync_bn(allreduce, world, peers=peers)
        tr.set_gradient_buckets(world > 1)
        for _ in range(3):
            tr.siamese_step(x1, x2, yd, world=world)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            lv, _ = tr.siamese_step(x1, x2, yd, world=world)
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = parallel.max_over_ranks(e0.elapsed_time(e1) / steps, dev)
        # end to end: host numpy batches -> pinned staging -> H2D -> step -> loss read back, every step
        for _ in range(2):
            tr.siamese_step(x1h, x2h, y, world=world)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            lv, _ = tr.siamese_step(x1h, x2h, y, world=world)
            loss_host = float(lv.item())
        t_e2e = parallel.max_over_ranks((time.perf_counter() - t0) / steps * 1e3, dev)
        launches = sum(1 for k, pl in tr._plans.items() for fn, _, _ in pl.ops if fn is not None) + 2   # + Adam (2 kernels)
        out["modes"][name] = dict(
            ms_per_step=round(ms, 4), value=round(args.train_pairs / (ms * 1e-3), 1),
            audio_seconds_per_sec=round(args.train_pairs * 2 * CLIP_SECONDS / (ms * 1e-3), 1),
            e2e=dict(ms_per_step=round(t_e2e, 4), value=round(args.train_pairs / (t_e2e * 1e-3), 1), unit="pairs/s",
                     h2d_bytes_per_step=2 * pairs * length * 4 + pairs * 4, d2h_bytes_per_step=4),
            backward={1: "one-plane fp16 gradients and operands, 1 MMA per MAC, per-block power-of-two scaling; "
                         "gradient tolerance 2e-3 (tests/test_gpu_train.py)",
                      3: "two-plane fp16 gradients and operands, 3 MMAs per MAC; gradient tolerance 5e-4"}[bwd],
            forward="fp16 x 3 (train-mode loss within 1e-4 of the oracle)",
            c_abi_calls_per_step=launches, final_loss=loss_host,
            roofline=dict(bound="tensor", achieved=round(3 * flops_fwd / (ms * 1e-3) / 1e12 / world, 1),
                          peak=peaks["bf16_tflops"], unit="TFLOP/s per GPU",
                          frac=round(3 * flops_fwd / (ms * 1e-3) / 1e12 / world / peaks["bf16_tflops"], 4),
                          note="ALGORITHMIC flops: 3 x forward (fwd + dgrad + wgrad) of 2 x pairs clips, counted once; "
                               f"forward issues 3 MMAs per MAC, backward {bwd}"))
        del tr, sia, enc
        torch.cuda.empty_cache()
    out.update(ms_per_step=out["modes"]["bwd1"]["ms_per_step"], value=out["modes"]["bwd1"]["value"], mode="bwd1",
               n_gpus=world, steps=steps, nccl_bytes_per_step=(1023808 + 2) * 4 if world > 1 else 0)
    return out


# ------------------------------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    from oracle import voicemap_oracle as O  # weights / synthetic input generators only (not timed, not compute)
    from voicemap_b200.engine import EncoderEngine
    from voicemap_b200.models import get_baseline_convolutional_encoder

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    n, length, steps, warmup = args.batch, args.length, args.steps, max(args.warmup, 3)
    params = O.init_encoder_params(FILTERS, EMB, seed=0, randomize_bn=True, random_bias=True)
    model = get_baseline_convolutional_encoder(FILTERS, EMB, input_shape=(length, 1))
    model.set_named_weights(params)
    model.precision = args.precision
    eng = model._get_engine()
    eng.pack()

    # inputs: rotate over a set larger than L2 (126 MB) so no step finds its input cached from the previous one
    n_sets = max(2, int(np.ceil(160e6 / (n * length * 4))))
    g = torch.Generator(device="cpu").manual_seed(1234 + rank)
    host_sets = [(O.WHITEN_RMS * torch.randn(n, length, generator=g)).pin_memory() for _ in range(n_sets)]
    dev_sets = [h.to(dev) for h in host_sets]
    out = torch.empty((n, EMB), dtype=torch.float32, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ("value")
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    for i in range(warmup):
        eng.forward(dev_sets[i % n_sets], out=out)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_w0 = time.perf_counter()
    e0.record()
    for i in range(steps):
        eng.forward(dev_sets[i % n_sets], out=out)
    e1.record()
    barrier()
    t_w1 = time.perf_counter()
    ms_total = e0.elapsed_time(e1)
    if sampler:
        sampler.window = (t_w0, t_w1)
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / steps
    value = world * n * CLIP_SECONDS / (ms_step * 1e-3)

    # ---- end to end through the public API: pinned host batch -> H2D -> encoder -> D2H embeddings
    for i in range(warmup):
        model.predict(host_sets[i % n_sets])
    barrier()
    t0 = time.perf_counter()
    for i in range(steps):
        emb_host = model.predict(host_sets[i % n_sets])
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e_value = world * n * CLIP_SECONDS * steps / e2e_s
    assert emb_host.shape == (n, EMB) and np.isfinite(emb_host).all()

    # ---- the reference's actual data contract: model.predict(numpy float64 (N, L, 1)) (voicemap/utils.py:133,156;
    # the batcher hands out float64): cast on host worker threads into pinned chunks + H2D + kernels + D2H
    host_f64 = [h.numpy().astype(np.float64)[:, :, None] for h in host_sets[:2]]
    for i in range(3):
        model.predict(host_f64[i % 2])
    barrier()
    f64_steps = min(steps, 40)
    t0 = time.perf_counter()
    for i in range(f64_steps):
        emb64 = model.predict(host_f64[i % 2])
    torch.cuda.synchronize()
    f64_s = time.perf_counter() - t0
    t = torch.tensor([f64_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_f64_value = world * n * CLIP_SECONDS * f64_steps / float(t.item())
    assert emb64.shape == (n, EMB) and np.isfinite(emb64).all()
    del host_f64

    # ---- sustained: the same device-resident step back to back for >= 2 s (clocks and power cap recorded)
    sus_steps = int(max(steps, np.ceil(args.sustained_seconds * 1e3 / ms_step)))
    sampler2 = ClockSampler(local_rank) if rank == 0 else None
    if sampler2:
        sampler2.start()
    barrier()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_s0 = time.perf_counter()
    s0.record()
    for i in range(sus_steps):
        eng.forward(dev_sets[i % n_sets], out=out)
    s1.record()
    barrier()
    t_s1 = time.perf_counter()
    if sampler2:
        sampler2.window = (t_s0, t_s1)
    sus_clocks = sampler2.stop() if sampler2 else None
    t = torch.tensor([s0.elapsed_time(s1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    sus_ms_step = float(t.item()) / sus_steps

    # ---- TF32 tensor peak, measured (the BASELINE.md bound was quoted against an assumed bf16 / 2)
    tf32_tflops = None
    if rank == 0:
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True
        ma = torch.randn(8192, 8192, device=dev)
        mb = torch.randn(8192, 8192, device=dev)
        best = 1e9
        for _ in range(6):
            m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            m0.record()
            torch.matmul(ma, mb)
            m1.record()
            torch.cuda.synchronize()
            best = min(best, m0.elapsed_time(m1))
        tf32_tflops = 2 * 8192 ** 3 / (best * 1e-3) / 1e12
        torch.backends.cuda.matmul.allow_tf32 = prev
        del ma, mb

    # ---- per-kernel durations (CUDA events on the launching stream) for the roofline object
    names = ["conv1", "conv3_b2", "conv3_b3", "conv3_b4", "gmax_dense"]
    prof_steps = min(steps, 20)
    all_evs = []
    # caller-owned activation planes, allocated once: no allocator work between the events
    x0 = dev_sets[0]
    h1 = eng.block1(x0)
    h2 = eng.block3(2, *h1)
    h3 = eng.block3(3, *h2)
    part = eng.block3(4, *h3, gmax=True)
    torch.cuda.synchronize()
    for i in range(prof_steps + 2):
        x = dev_sets[i % n_sets]
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        evs[0].record()
        eng.block1(x, out=h1); evs[1].record()
        eng.block3(2, *h1, out=h2); evs[2].record()
        eng.block3(3, *h2, out=h3); evs[3].record()
        eng.block3(4, *h3, gmax=True, out=part); evs[4].record()
        eng.gmax_dense(part); evs[5].record()
        all_evs.append(evs)
    torch.cuda.synchronize()   # one sync at the end: the stream never drains between kernels
    acc = {k: 0.0 for k in names}
    for evs in all_evs[2:]:
        for j, k in enumerate(names):
            acc[k] += evs[j].elapsed_time(evs[j + 1])
    kern_ms = {k: v / prof_steps for k, v in acc.items()}

    # ---- the same clips without the reference's x4 decimation (raw 16 kHz: 48000 samples per 3 s clip), for the record
    raw16k = None
    if length != 48000:
        xr = O.WHITEN_RMS * torch.randn(64, 48000, generator=torch.Generator(device="cpu").manual_seed(99))
        xr = xr.to(dev)
        outr = torch.empty((64, EMB), dtype=torch.float32, device=dev)
        for _ in range(3):
            eng.forward(xr, out=outr)
        torch.cuda.synchronize()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record()
        for _ in range(20):
            eng.forward(xr, out=outr)
        r1.record()
        torch.cuda.synchronize()
        raw_ms = r0.elapsed_time(r1) / 20
        raw16k = dict(value=round(world * 64 * CLIP_SECONDS / (raw_ms * 1e-3), 1), unit=UNIT, batch=64, length=48000,
                      ms_per_step=round(raw_ms, 4),
                      note="same encoder on un-decimated 16 kHz clips (the reference's scripts decimate x4 first, "
                           "voicemap/utils.py:29; SURVEY.md F3); rank 0's timing x n_gpus, informational")
        del xr, outr

    peaks = load_peaks()
    train = None if args.no_train else bench_train(args, dev, rank, world, peaks)

    if rank == 0:
        work = block_work(length)
        conv3_ms = kern_ms["conv3_b2"] + kern_ms["conv3_b3"] + kern_ms["conv3_b4"]
        conv3_flop = sum(b["flop"] for b in work[1:]) * n
        achieved_tf = conv3_flop / (conv3_ms * 1e-3) / 1e12
        blocks = []
        for b, k in zip(work, names[:4]):
            t_ms = kern_ms[k]
            blocks.append(dict(kernel=k, ms=round(t_ms, 4),
                               tflops=round(b["flop"] * n / (t_ms * 1e-3) / 1e12, 1),
                               gbs=round(b["bytes"] * n / (t_ms * 1e-3) / 1e9, 1)))
        # network-level bound of BASELINE.md: sum_blocks max(t_HBM, t_tensor); tensor peak named explicitly
        def bound_us(tensor_tflops):
            return sum(max(b["bytes"] / (peaks["hbm_gbs"] * 1e9), b["flop"] / (tensor_tflops * 1e12)) for b in work) * 1e6
        us_per_clip = ms_step * 1e3 / n
        roofline = dict(
            bound="tensor", kernel="conv3_kernel (blocks 2-4, 3 launches/step)",
            achieved=round(achieved_tf, 1), peak=peaks["bf16_tflops"], unit="TFLOP/s",
            frac=round(achieved_tf / peaks["bf16_tflops"], 4),
            peak_source=peaks["source"] + ", burst bf16 figure",
            note="achieved counts ALGORITHMIC flops 2*L*K*Cin*Cout once; " + {
                3: "precision=3 issues 3 fp16 MMAs per algorithmic MAC (fp32-grade split), so tensor-pipe time is 3x "
                   "this fraction",
                2: "precision=2 issues one fp16 MMA plus one fp8 MMA over a doubled K (both correction products as "
                   "e5m2 byte pairs, fp8 runs at twice the fp16 rate) per algorithmic MAC, so tensor-pipe time is 2x "
                   "this fraction",
                1: "precision=1: one fp16 MMA per algorithmic MAC (not a parity mode)"}[args.precision],
            mma_issue_frac=round(achieved_tf * args.precision / peaks["bf16_tflops"], 4),
            # dram__bytes_read + dram__bytes_write of the three conv3 launches (ncu --set full, profiles/r02b_ncu_summary.md:
            # 735 + 652 + 306 MB at batch 256, the same for precision 2 and 3: both move 4 bytes per activation)
            # averaged per launch; algorithmic bytes per launch average 590 MB
            traffic=(5.64e8 if (n, length) == (256, 12000) and args.precision in (2, 3) else None),
            traffic_source="ncu --set full capture of the same kernels at this batch, profiles/r02b_ncu_summary.md "
                           "(dram__bytes_read.sum + dram__bytes_write.sum, mean of the three conv3 launches)",
            blocks=blocks,
            network=dict(us_per_clip=round(us_per_clip, 3),
                         frac_of_bf16_roofline=round(bound_us(peaks["bf16_tflops"]) / us_per_clip, 4),
                         tf32_tflops_measured=round(tf32_tflops, 1),
                         frac_of_tf32_roofline=round(bound_us(tf32_tflops) / us_per_clip, 4),
                         note="BASELINE.md section 3 bounds: sum over blocks of max(bytes/HBM, flops/peak); the TF32 "
                              "peak is measured in this run (torch.matmul fp32 8192^3 with TF32 allowed, best of 6)"))
        cpu_baseline = None
        if not args.no_cpu_baseline and world == 1:     # the CPU leg is timed at N = 1 only (rank 0's host cores, alone)
            cpu_baseline, _ = time_oracle(length, 8, budget_s=args.cpu_baseline_seconds)
        line = dict(
            metric=METRIC, value=round(value, 1), unit=UNIT, n_gpus=world, steps=steps, warmup=warmup,
            ms_per_step=round(ms_step, 4), higher_is_better=True, scaling="weak", vs_baseline=None,
            dtype={3: "fp16x3->f32 (fp32-grade split on fp16 tensor cores)",
                   2: "fp16+fp8(e5m2 pairs)->f32 (split precision: fp16 main product, fp8 correction product, fp32 accumulate)",
                   1: "fp16 (fp32 accumulate)"}[args.precision],
            data="synthetic",
            config=dict(workload=f"baseline 1D-CNN encoder fwd (eval), filters={FILTERS}, emb={EMB}, batch={n} clips/GPU, "
                                 f"{length} samples/clip (3 s @ 16 kHz {'raw' if length == 48000 else 'after the reference x4 decimation'}), "
                                 f"precision={args.precision}",
                        l2=f"inputs rotate over {n_sets} batches ({n_sets * n * length * 4 / 1e6:.0f} MB > 126 MB L2); "
                           f"activations {eng.lib.vm_encoder_workspace_bytes(n, length, FILTERS, 4) / 1e6:.0f} MB/step",
                        parallelism=f"dp{world} (independent clips, no collective)"),
            clocks=clocks,
            e2e=dict(value=round(e2e_value, 1), unit=UNIT, h2d_bytes_per_step=n * length * 4,
                     d2h_bytes_per_step=n * EMB * 4,
                     note="model.predict(pinned host float32 batch): H2D + 5 kernels + D2H + sync, wall clock"),
            e2e_numpy_f64=dict(value=round(e2e_f64_value, 1), unit=UNIT, host_bytes_per_step=n * length * 8,
                               h2d_bytes_per_step=n * length * 4, d2h_bytes_per_step=n * EMB * 4, steps=f64_steps,
                               note="model.predict(numpy float64 (N, L, 1)), the reference's own call "
                                    "(voicemap/utils.py:133,156): float64 -> float32 cast on host worker threads into "
                                    "pinned chunks, H2D, kernels, D2H, wall clock"),
            sustained=dict(value=round(world * n * CLIP_SECONDS / (sus_ms_step * 1e-3), 1), unit=UNIT,
                           ms_per_step=round(sus_ms_step, 4), steps=sus_steps,
                           seconds=round(sus_ms_step * sus_steps / 1e3, 2), clocks=sus_clocks,
                           conv3_frac_of_bf16_sustained=round(
                               achieved_tf * (conv3_ms / (conv3_ms + kern_ms["conv1"] + kern_ms["gmax_dense"]))
                               * (ms_step / sus_ms_step) / peaks["bf16_tflops_sustained"], 4),
                           note="device-resident steps back to back for >= --sustained-seconds; the fraction scales "
                                "the conv3 roofline figure by the sustained / burst step-time ratio and compares it "
                                "with the SUSTAINED measured bf16 peak"),
            gpu_launches=5 * steps,
            kernel_ms=kern_ms,
            raw16k=raw16k,
            roofline=roofline,
            cpu_baseline=cpu_baseline,
            train=train,
        )
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
