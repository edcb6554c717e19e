"""N>1 host logic on CPU: two gloo ranks exercise sharding, the flat-gradient all-reduce (+ identical Adam update on
every rank), max-over-ranks timing and uneven-shard loss means."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import voicemap_oracle as O
from voicemap_b200 import parallel


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world_size, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world_size))
    parallel.init_from_env(backend="gloo")
    assert parallel.world() == (rank, world_size)
    # sharding: 7 pairs over 2 ranks -> [0,4) and [4,7)
    lo, hi = parallel.shard_bounds(7)
    # per-rank "gradients" of a toy quadratic on this rank's shard; the all-reduced sum must equal the full-batch
    # gradient and the Adam update must be bit-identical on both ranks
    rng = np.random.default_rng(0)
    data = rng.normal(size=(7, 5))
    w = np.linspace(-1, 1, 5)
    local = torch.from_numpy((data[lo:hi] * (data[lo:hi] @ w)[:, None]).sum(axis=0))
    flat = parallel.allreduce_sum_(local.clone())
    full = (data * (data @ w)[:, None]).sum(axis=0)
    p = {"w": w.copy()}
    O.keras_adam_step(p, {"w": flat.numpy() / 7.0}, {"w": np.zeros(5)}, {"w": np.zeros(5)}, t=1, clipnorm=1.0)
    # bucketed asynchronous all-reduce of a flat buffer: same result as one all-reduce, buckets issued out of order
    flat2 = torch.arange(10, dtype=torch.float64) * (rank + 1)
    buckets = parallel.GradientBuckets(flat2, [(6, 10), (3, 6), (0, 3)])
    for i in range(3):
        buckets.launch(i)
    buckets.wait()
    assert torch.equal(flat2, torch.arange(10, dtype=torch.float64) * 3) and not buckets.works
    try:
        parallel.GradientBuckets(flat2, [(0, 3), (4, 10)])
        raise AssertionError("a gap between buckets must be rejected")
    except ValueError:
        pass
    t_max = parallel.max_over_ranks(1.0 + rank)
    mean = parallel.global_mean(float(np.arange(lo, hi).sum()), hi - lo)
    rows = parallel.gather_rows(torch.full((hi - lo, 2), float(rank)))
    # the reference's builders take no seed: ranks start from different weights until rank 0's are broadcast
    from voicemap_b200.models import build_siamese_net, get_baseline_convolutional_encoder
    model = build_siamese_net(get_baseline_convolutional_encoder(16, 8), (1024, 1))
    before = np.concatenate([w.reshape(-1) for w in model.get_weights()])
    parallel.broadcast_weights_(model)
    after = np.concatenate([w.reshape(-1) for w in model.get_weights()])
    # the n-shot callback pools its accuracy over ranks: rank 0 solves every task, rank 1 none -> 0.5 on both
    from voicemap_b200 import utils

    class Tasks:
        def build_n_shot_task(self, k, n=1):
            query = np.full(64, 1.0)
            support = np.stack([np.full(64, 1.0 if (c == 0) == (rank == 0) else 5.0) for c in range(k) for _ in range(n)])
            return (query, 0), (support, np.repeat(np.arange(k), n))

    class Siamese:
        layers = [None, None, None]

        def predict(self, x):
            return np.abs(x[0] - x[1]).mean(axis=(1, 2))[:, None]

    callback = utils.NShotEvaluationCallback(6, 1, 3, Tasks(), preprocessor=utils.BatchPreProcessor(
        "siamese", utils.preprocess_instances(1, whitening=False)))
    callback.set_model(Siamese())
    logs = {}
    callback.on_epoch_end(0, logs)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), lo=lo, hi=hi, flat=flat.numpy(), full=full, w=p["w"], t_max=t_max,
             mean=mean, rows=rows.numpy(), before=before, after=after,
             pooled=logs["val_1-shot_acc"])
    dist.destroy_process_group()


def test_two_rank_gloo(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "r0.npz"), np.load(tmp_path / "r1.npz")
    assert (int(r0["lo"]), int(r0["hi"]), int(r1["lo"]), int(r1["hi"])) == (0, 4, 4, 7)
    np.testing.assert_allclose(r0["flat"], r0["full"], rtol=1e-12)
    np.testing.assert_array_equal(r0["flat"], r1["flat"])
    np.testing.assert_array_equal(r0["w"], r1["w"])          # identical update on every rank
    assert r0["t_max"] == r1["t_max"] == 2.0                  # slowest rank defines the step time
    assert abs(float(r0["mean"]) - 3.0) < 1e-12               # mean over the global batch, uneven shards
    assert r0["rows"].shape == (7, 2) and r0["rows"][:4].sum() == 0 and r0["rows"][4:].sum() == 6
    assert not np.array_equal(r0["before"], r1["before"])     # independently initialised ranks differ ...
    np.testing.assert_array_equal(r0["after"], r0["before"])  # ... rank 0 keeps its weights ...
    np.testing.assert_array_equal(r1["after"], r0["before"])  # ... and rank 1 receives them, bit for bit
    assert float(r0["pooled"]) == float(r1["pooled"]) == 0.5   # pooled n-shot accuracy, identical on both ranks


def test_single_process_defaults():
    assert parallel.world() == (0, 1)
    assert parallel.shard_bounds(10) == (0, 10)
    assert parallel.shard_bounds(10, 1, 4) == (3, 6) and parallel.shard_bounds(10, 3, 4) == (8, 10)
    t = torch.ones(3)
    assert parallel.allreduce_sum_(t) is t and parallel.max_over_ranks(2.5) == 2.5
    marker = object()
    assert parallel.broadcast_weights_(marker) is marker       # one process: nothing to do
