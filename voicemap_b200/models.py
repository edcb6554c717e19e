"""voicemap.models on B200: the two builder functions of the reference with unchanged signatures
(voicemap/models.py:6 and :44), returning objects that expose the Keras-method subset voicemap's scripts call
(SURVEY.md 8(b)).  All network arithmetic runs in libvoicemap_b200.so through ``EncoderEngine``.
"""
from __future__ import annotations

import json
import os
from collections import OrderedDict
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from .keras_compat import Adam, Dense

DISTANCE_METRICS = ('uniform_euclidean', 'weighted_euclidean',
                    'uniform_l1', 'weighted_l1',
                    'dot_product', 'cosine_distance')

_PREDICT_CHUNK = 4096  # clips per device launch in predict(); results do not depend on it (eval mode)
_PIPELINE_BUFFER_BYTES = 1 << 31  # device staging buffer for host batches in predict()


def _pipeline_plan(n, pinned=True):
    """[(lo, hi), ...] covering range(n).  A pinned batch is cut into a small first chunk (its copy is the only one
    that is not hidden behind kernels) and the rest, in pieces of at most _PREDICT_CHUNK clips.  The first chunk is
    3/16 of the batch: with the precision-2 kernels a 256-clip batch computes in ~0.83 ms and copies in ~0.23 ms, so
    the second copy (13/16 of the bytes) just fits behind the first chunk's kernels plus their fixed ~0.06 ms of
    launches and prologues.  Every extra chunk costs that fixed part again, so finer plans are not faster
    (tools/e2e_probe.py, profiles/r01_e2e_probe.log).  Pageable memory copies synchronously: plain chunks."""
    if not pinned or n < 96:
        return [(lo, min(n, lo + _PREDICT_CHUNK)) for lo in range(0, n, _PREDICT_CHUNK)]
    first = min(_PREDICT_CHUNK, max(16, n * 3 // 16 // 8 * 8))
    plan = [(0, first)]
    lo = first
    while lo < n:
        hi = min(n, lo + _PREDICT_CHUNK)
        plan.append((lo, hi))
        lo = hi
    return plan


def _cast_plan(n):
    """[(lo, hi), ...]: staging chunks of a numpy batch of n clips -- half of the batch each, 16 .. 128 clips.  Every
    chunk costs the encoder its fixed ~0.06 ms of launches and prologues, so chunks are large; a chunk is cast as row
    blocks on all of ``_HostStage``'s threads, so the first one is ready after half of the whole cast.  Measured on the
    B200 box for 256 x 12000 doubles (tools/predict_numpy_bench.py, profiles/r02_predict_numpy_staging_probe.log;
    predict() of the same batch as pinned float32: 1.03 ms): one thread per chunk, 4 chunks of 64 clips in parallel (the
    first form) 1.65 - 1.75 ms; row blocks on 3 threads: 2 x 128 clips 1.45 - 1.48 ms, 3 x 86 1.47 - 1.49, 4 x 64
    1.51 - 1.53, 64 + 192 1.65 - 1.74, 1 x 256 1.92 - 1.96; the same on 2 / 4 / 6 threads 1.48 - 1.53 / 1.55 - 1.56 /
    1.72 - 1.73 ms (more threads take the GIL from the thread that issues the launches)."""
    rows = int(min(n, 128, max(16, -(-n // 2))))
    return [(lo, min(n, lo + rows)) for lo in range(0, n, rows)]


class _HostStage:
    """numpy batches (the batcher and the reference's preprocessing hand out float64, voicemap/librispeech.py:103) ->
    float32 chunks in pinned memory.  The cast is the expensive part of ``predict(numpy)`` -- 1.5 ms (B200 box) to 15 ms
    (build container) on one host thread for 256 x 12000 doubles, against 0.9 ms of kernels -- so it runs chunk by chunk on worker threads (numpy releases the
    GIL while it converts; a chunk = row blocks on all threads), a few chunks ahead of the copy engine, and every chunk
    is copied and embedded as soon as it is ready.  Slots are pinned once and reused; a slot is recycled after the copy that read it has finished."""

    def __init__(self):
        try:
            cpus = len(os.sched_getaffinity(0))
        except AttributeError:
            cpus = os.cpu_count() or 1
        self.workers = max(1, min(3, cpus))
        self.pieces = True       # a chunk is cast as row blocks on all threads (False: one thread per chunk)
        self.pool = ThreadPoolExecutor(max_workers=self.workers)
        self.slots = []          # [pinned float32 tensor, copy-done event or None]

    def _slot(self, k, elems):
        import torch
        while len(self.slots) <= k:
            self.slots.append([None, None])
        slot = self.slots[k]
        if slot[1] is not None:
            slot[1].synchronize()
            slot[1] = None
        if slot[0] is None or slot[0].numel() < elems:
            slot[0] = torch.empty(elems, dtype=torch.float32).pin_memory()
        return slot

    def chunks(self, src):
        """src: numpy (n, length) view of any dtype, any strides.  Yields (lo, hi, pinned float32 (hi-lo, length), slot)
        in order; the consumer stores the event of its copy in ``slot[1]``."""
        n, length = src.shape
        plan = _cast_plan(n)
        count = len(plan)
        rows = max(hi - lo for lo, hi in plan)
        depth = min(count, 4)
        pending = {}

        pieces = self.workers if self.pieces else 1

        def submit(i):
            lo, hi = plan[i]
            slot = self._slot(i % depth, rows * length)
            view = slot[0][:(hi - lo) * length].view(hi - lo, length)
            dst = view.numpy()
            # a chunk goes to the pool as `pieces` row blocks: the pool is first-in first-out, so all threads work on
            # the oldest chunk and chunk 0 is ready after 1/count of the whole cast instead of all chunks finishing
            # together at the end of it
            step = -(-(hi - lo) // pieces)
            futs = [self.pool.submit(np.copyto, dst[r:r + step], src[lo + r:min(hi, lo + r + step)], casting='unsafe')
                    for r in range(0, hi - lo, step)]
            pending[i] = (futs, lo, hi, view, slot)

        for i in range(depth):
            submit(i)
        for i in range(count):
            futs, lo, hi, view, slot = pending.pop(i)
            for f in futs:
                f.result()
            yield lo, hi, view, slot
            if i + depth < count:
                submit(i + depth)


def _glorot_uniform(shape, rng):
    if len(shape) == 2:
        fan_in, fan_out = shape
    else:
        rec = int(np.prod(shape[:-2]))
        fan_in, fan_out = shape[-2] * rec, shape[-1] * rec
    limit = np.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-limit, limit, size=shape).astype(np.float32)


class LayerInfo:
    """Entry of ``model.layers`` (name / class / config / weight names), enough for summary() and indexing."""

    def __init__(self, name, class_name, config=None, weight_names=(), model=None):
        self.name = name
        self.class_name = class_name
        self.config = dict(config or {})
        self.weight_names = tuple(weight_names)
        self._model = model

    def __repr__(self):
        return f"<{self.class_name} {self.name}>"


class _ModelBase:
    optimizer = None
    loss = None
    metrics = None
    _engine = None

    # ---- Keras plumbing shared by both model kinds
    def compile(self, loss=None, optimizer=None, metrics=None, **_):
        """model.compile(loss=<str|callable>, optimizer=Adam(...), metrics=['accuracy'])
        (experiments/train_siamese.py:57, siamese_contrastive_loss.py:70)."""
        from .utils import contrastive_loss
        if callable(loss):
            if loss is not contrastive_loss:
                raise NotImplementedError("only voicemap.utils.contrastive_loss is supported as a callable loss")
            loss = "contrastive_loss"
        if loss not in ("binary_crossentropy", "categorical_crossentropy", "contrastive_loss"):
            raise NotImplementedError(f"loss {loss!r} is not used on the voicemap path")
        self.loss = loss
        self.optimizer = optimizer if optimizer is not None else Adam()
        self.metrics = list(metrics or [])

    def fit_generator(self, generator, steps_per_epoch=None, epochs=1, verbose=1, callbacks=None,
                      validation_data=None, validation_steps=None, workers=1, use_multiprocessing=False,
                      initial_epoch=0, **_):
        from .training import fit_generator
        return fit_generator(self, generator, steps_per_epoch=steps_per_epoch, epochs=epochs, verbose=verbose,
                             callbacks=callbacks, validation_data=validation_data,
                             validation_steps=validation_steps, initial_epoch=initial_epoch, workers=workers,
                             use_multiprocessing=use_multiprocessing)

    def _non_trainable_params(self):
        encoder = getattr(self, "encoder", self)
        return int(sum(w.size for name, w in encoder.weights.items() if name.endswith(("_mean", "_var"))))

    def summary(self, print_fn=None):
        lines = ["_" * 65, f"{'Layer (type)':<40}{'Param #':>25}", "=" * 65]
        total = 0
        for layer in self.layers:
            if isinstance(layer, _ModelBase):     # the shared encoder inside a siamese model
                n = layer.count_params()
                lines.append(f"{layer.name + ' (Sequential)':<40}{n:>25}")
            else:
                n = int(sum(np.prod(self._weight(w).shape) for w in layer.weight_names))
                lines.append(f"{layer.name + ' (' + layer.class_name + ')':<40}{n:>25}")
            total += n
        frozen = self._non_trainable_params()     # BatchNormalization moving statistics
        lines += ["=" * 65, f"Total params: {total:,}", f"Trainable params: {total - frozen:,}",
                  f"Non-trainable params: {frozen:,}", "_" * 65]
        text = "\n".join(lines)
        (print_fn or print)(text)
        return None

    def save(self, filepath, overwrite=True):
        """model.save (keras ModelCheckpoint call site: experiments/train_siamese.py:81-87).  ``*.h5`` / ``*.hdf5``
        paths get a Keras-2.2.x-layout HDF5 checkpoint (voicemap_b200/keras_hdf5.py: model_config, training_config,
        /model_weights/<layer>/<weight>), any other path this package's npz container.  ``load_model`` reads both."""
        if str(filepath).lower().endswith((".h5", ".hdf5")):
            from .keras_hdf5 import save_keras_weights
            training = None
            if self.loss is not None:
                opt = self.optimizer
                training = dict(loss=self.loss, metrics=list(self.metrics or []), loss_weights=None,
                                sample_weight_mode=None,
                                optimizer_config=dict(class_name="Adam", config=opt.get_config()) if opt else None)
            save_keras_weights(filepath, self._keras_config(), self._keras_layers(), training)
            return
        arrays = {f"w{i}": w for i, w in enumerate(self.get_weights())}
        arrays["config"] = np.frombuffer(json.dumps(self.get_config()).encode(), dtype=np.uint8)
        with open(filepath, "wb") as fh:
            np.savez(fh, **arrays)


_GLOROT = {"class_name": "VarianceScaling",
           "config": {"distribution": "uniform", "scale": 1.0, "seed": None, "mode": "fan_avg"}}
_ZEROS, _ONES = {"class_name": "Zeros", "config": {}}, {"class_name": "Ones", "config": {}}


def _dense_config(name, units, activation):
    return {"class_name": "Dense", "config": {
        "name": name, "trainable": True, "units": int(units), "activation": activation, "use_bias": True,
        "kernel_initializer": _GLOROT, "bias_initializer": _ZEROS, "kernel_regularizer": None,
        "bias_regularizer": None, "activity_regularizer": None, "kernel_constraint": None, "bias_constraint": None}}


class EncoderModel(_ModelBase):
    """Sequential returned by get_baseline_convolutional_encoder (voicemap/models.py:6-41), optionally extended
    with a classification head through ``add(Dense(n, activation='softmax'))``."""

    def __init__(self, filters, embedding_dimension, input_shape=None, dropout=0.05, seed=None, name="sequential_1",
                 first_pool=4):
        """first_pool = 4 is voicemap/models.py:19; first_pool = 2 rebuilds the older architecture of the checkpoint
        shipped under models/n_seconds/ (four MaxPooling1D(2)), which ``load_model`` selects itself."""
        if first_pool not in (2, 4):
            raise ValueError("first_pool must be 4 (voicemap/models.py:19) or 2 (older checkpoints)")
        self.first_pool = int(first_pool)
        self.name = name
        self.filters = int(filters)
        self.embedding_dimension = int(embedding_dimension)
        self.input_shape = tuple(input_shape) if input_shape is not None else None
        self.dropout = float(dropout)
        self._precision = 2
        rng = np.random.default_rng(seed)
        f = self.filters
        self.weights = OrderedDict()
        self.layers = []
        cin = 1
        for i, (k, mult, pool) in enumerate(((32, 1, self.first_pool), (3, 2, 2), (3, 3, 2), (3, 4, 2)), start=1):
            cout = mult * f
            self.weights[f"conv{i}_kernel"] = _glorot_uniform((k, cin, cout), rng)
            self.weights[f"conv{i}_bias"] = np.zeros((cout,), np.float32)
            self.weights[f"bn{i}_gamma"] = np.ones((cout,), np.float32)
            self.weights[f"bn{i}_beta"] = np.zeros((cout,), np.float32)
            self.weights[f"bn{i}_mean"] = np.zeros((cout,), np.float32)
            self.weights[f"bn{i}_var"] = np.ones((cout,), np.float32)
            self.layers += [
                LayerInfo(f"conv1d_{i}", "Conv1D", dict(filters=cout, kernel_size=k, padding="same", activation="relu"),
                          (f"conv{i}_kernel", f"conv{i}_bias"), self),
                LayerInfo(f"batch_normalization_{i}", "BatchNormalization", dict(epsilon=1e-3, momentum=0.99),
                          (f"bn{i}_gamma", f"bn{i}_beta", f"bn{i}_mean", f"bn{i}_var"), self),
                LayerInfo(f"spatial_dropout1d_{i}", "SpatialDropout1D", dict(rate=self.dropout), (), self),
                LayerInfo(f"max_pooling1d_{i}", "MaxPooling1D", dict(pool_size=pool, strides=pool), (), self),
            ]
            cin = cout
        self.layers.append(LayerInfo("global_max_pooling1d_1", "GlobalMaxPooling1D", {}, (), self))
        self.weights["dense_kernel"] = _glorot_uniform((4 * f, self.embedding_dimension), rng)
        self.weights["dense_bias"] = np.zeros((self.embedding_dimension,), np.float32)
        self.layers.append(LayerInfo("dense_1", "Dense", dict(units=self.embedding_dimension, activation="linear"),
                                     ("dense_kernel", "dense_bias"), self))
        self._n_encoder_layers = len(self.layers)
        self._head = None          # classification head: dict(units, activation)
        self._has_embedding_dense = True
        self._rng = rng
        self._engine = None
        self._engine_dirty = True

    # ---- Sequential API
    def add(self, layer):
        """classifier.add(Dense(num_classes, activation='softmax')) (experiments/train_classifier.py:112)."""
        if not isinstance(layer, Dense):
            raise NotImplementedError("only a Dense head can be added to the voicemap encoder")
        if self._head is not None:
            raise NotImplementedError("a single Dense head is supported")
        self._head = dict(units=layer.units, activation=layer.activation or "linear")
        self.weights["head_kernel"] = _glorot_uniform((self.embedding_dimension, layer.units), self._rng)
        self.weights["head_bias"] = np.zeros((layer.units,), np.float32)
        self.layers.append(LayerInfo("dense_2", "Dense", dict(units=layer.units, activation=layer.activation),
                                     ("head_kernel", "head_bias"), self))

    def pop(self):
        """Remove the last layer (voicemap/utils.py:145 strips the softmax head of a classifier clone)."""
        if self._head is None:
            raise NotImplementedError("only an added Dense head can be popped from the voicemap encoder")
        self.layers.pop()
        self._head = None
        del self.weights["head_kernel"], self.weights["head_bias"]

    def _weight(self, name):
        return self.weights[name]

    def count_params(self):
        return int(sum(w.size for w in self.weights.values()))

    def get_weights(self):
        """Keras order: per layer kernel,bias / gamma,beta,moving_mean,moving_variance / ... / dense / head."""
        return [w.copy() for w in self.weights.values()]

    def set_weights(self, weights):
        weights = list(weights)
        if len(weights) != len(self.weights):
            raise ValueError(f"You called `set_weights(weights)` with a weight list of length {len(weights)}, "
                             f"but the model was expecting {len(self.weights)} weights.")
        for (name, old), new in zip(self.weights.items(), weights):
            new = np.asarray(new, dtype=np.float32)
            if new.shape != old.shape:
                raise ValueError(f"Layer weight shape {old.shape} not compatible with provided weight shape "
                                 f"{new.shape} ({name})")
            self.weights[name] = new.copy()
        self._engine_dirty = True
        self._drop_trainer()

    def _drop_trainer(self):
        """Weights set from the host supersede the device copies a TrainEngine holds (its parameter buffer, Adam
        moments and recorded launch plans): the next fit / train step builds a fresh one from the new weights."""
        for owner in (self, getattr(self, "_siamese_owner", None)):
            if owner is not None and getattr(owner, "_trainer", None) is not None:
                owner._trainer = None

    def set_named_weights(self, mapping):
        for name, value in mapping.items():
            value = np.asarray(value, dtype=np.float32)
            if value.shape != self.weights[name].shape:
                raise ValueError(f"{name}: expected {self.weights[name].shape}, got {value.shape}")
            self.weights[name] = value.copy()
        self._engine_dirty = True
        self._drop_trainer()

    def get_config(self):
        return dict(kind="encoder", filters=self.filters, embedding_dimension=self.embedding_dimension,
                    input_shape=self.input_shape, dropout=self.dropout, head=self._head, first_pool=self.first_pool)

    @property
    def precision(self):
        """Arithmetic of the eval forward: 2 (default) = fp16 product + one fp8 (e5m2 pairs)
        correction product, embeddings 1e-5 .. 4.5e-5 from the fp64 oracle; 3 = fp16 x 3, ~7e-6; 1 = fp16 x 1, ~6e-4
        (throughput mode, not parity).  Training always runs 3."""
        return self._precision

    @precision.setter
    def precision(self, value):
        if int(value) not in (1, 2, 3):
            raise ValueError("precision must be 1, 2 or 3")
        if int(value) != self._precision:
            self._precision = int(value)
            if self._engine is not None:     # the packed weights carry the planes of every mode: the arithmetic is a
                self._engine.precision = int(value)   # per-launch argument, so the engine (and a trainer's tensors
                                                      # that live in it) stays

    def _clone(self):
        m = EncoderModel(self.filters, self.embedding_dimension, self.input_shape, self.dropout,
                         first_pool=self.first_pool)
        m.precision = self.precision
        if self._head is not None:
            m.add(Dense(self._head["units"], activation=self._head["activation"]))
        return m

    # ---- Keras 2.2.x serialisation (what keras/engine/saving.py stores; keys as in the reference's checkpoints)
    def _keras_layer_configs(self):
        shape = [None] + [int(v) if v is not None else None for v in (self.input_shape or (None, 1))]
        out = []
        for i, (k, mult, pool) in enumerate(((32, 1, self.first_pool), (3, 2, 2), (3, 3, 2), (3, 4, 2)), start=1):
            conv = {"name": f"conv1d_{i}", "trainable": True, "filters": mult * self.filters, "kernel_size": [k],
                    "strides": [1], "padding": "same", "data_format": "channels_last", "dilation_rate": [1],
                    "activation": "relu", "use_bias": True, "kernel_initializer": _GLOROT,
                    "bias_initializer": _ZEROS, "kernel_regularizer": None, "bias_regularizer": None,
                    "activity_regularizer": None, "kernel_constraint": None, "bias_constraint": None}
            if i == 1:
                conv["batch_input_shape"] = shape
            out.append({"class_name": "Conv1D", "config": conv})
            out.append({"class_name": "BatchNormalization", "config": {
                "name": f"batch_normalization_{i}", "trainable": True, "axis": -1, "momentum": 0.99, "epsilon": 0.001,
                "center": True, "scale": True, "beta_initializer": _ZEROS, "gamma_initializer": _ONES,
                "moving_mean_initializer": _ZEROS, "moving_variance_initializer": _ONES, "beta_regularizer": None,
                "gamma_regularizer": None, "beta_constraint": None, "gamma_constraint": None}})
            out.append({"class_name": "SpatialDropout1D", "config": {
                "name": f"spatial_dropout1d_{i}", "trainable": True, "rate": self.dropout, "noise_shape": None,
                "seed": None}})
            out.append({"class_name": "MaxPooling1D", "config": {
                "name": f"max_pooling1d_{i}", "trainable": True, "pool_size": [pool], "strides": [pool],
                "padding": "valid"}})
        out.append({"class_name": "GlobalMaxPooling1D", "config": {"name": "global_max_pooling1d_1", "trainable": True}})
        out.append(_dense_config("dense_1", self.embedding_dimension, "linear"))
        if self._head is not None:
            out.append(_dense_config("dense_2", self._head["units"], self._head["activation"]))
        return out

    def _keras_config(self):
        return {"class_name": "Sequential", "config": self._keras_layer_configs()}   # Keras 2.2.2: a plain list

    _KERAS_NAMES = {"kernel": "kernel", "bias": "bias", "gamma": "gamma", "beta": "beta", "mean": "moving_mean",
                    "var": "moving_variance"}

    def _keras_weight_items(self, trainable_first=False):
        """[(keras layer name, keras weight name, array)] in Keras order."""
        items = []
        for key, arr in self.weights.items():
            blk, kind = key.rsplit("_", 1)
            if blk.startswith("conv"):
                layer = f"conv1d_{blk[4:]}"
            elif blk.startswith("bn"):
                layer = f"batch_normalization_{blk[2:]}"
            else:
                layer = "dense_1" if blk == "dense" else "dense_2"
            items.append((layer, f"{layer}/{self._KERAS_NAMES[kind]}:0", arr))
        if trainable_first:   # nested model: trainable weights, then the moving statistics
            moving = [it for it in items if "moving_" in it[1]]
            items = [it for it in items if "moving_" not in it[1]] + moving
        return items

    def _keras_layers(self):
        by_layer = OrderedDict((c["config"]["name"], []) for c in self._keras_layer_configs())
        for layer, wname, arr in self._keras_weight_items():
            by_layer[layer].append((wname, arr))
        return list(by_layer.items())

    # ---- device
    def _get_engine(self):
        from .engine import EncoderEngine
        if self._engine is None:
            self._engine = EncoderEngine(self.filters, self.embedding_dimension, precision=self.precision,
                                         first_pool=self.first_pool)
            self._engine_dirty = True
        if self._engine_dirty:
            self._engine.set_weights(self.weights)
            self._engine_dirty = False
        return self._engine

    def embed_device(self, x_dev):
        """x_dev: CUDA fp32 (N, L[, 1]) -> CUDA (N, embedding_dimension)."""
        return self._get_engine().forward(x_dev)

    def _check_input_shape(self, shape):
        if len(shape) != 3 or shape[2] != 1:
            raise ValueError(f"Error when checking input: expected input to have shape (N, L, 1) but got array "
                             f"with shape {shape}")
        if self.input_shape is not None and tuple(shape[1:]) != tuple(self.input_shape):
            raise ValueError(f"Error when checking input: expected conv1d_1_input to have shape "
                             f"{self.input_shape} but got array with shape {shape[1:]}")

    def _host_batch(self, x):
        """numpy (N, L, 1) of any float dtype, or a torch CPU tensor (N, L[, 1]) (pinned memory makes the
        host->device copy asynchronous) -> contiguous float32 torch CPU tensor (N, L)."""
        import torch
        if isinstance(x, torch.Tensor):
            if x.is_cuda:
                raise ValueError("predict() takes host data; use embed_device() for CUDA tensors")
            if x.dim() == 2:
                x = x.unsqueeze(-1)
            shape = tuple(x.shape)
        else:
            x = np.asarray(x)
            shape = x.shape
        self._check_input_shape(shape)
        if isinstance(x, torch.Tensor):
            xt = x.reshape(shape[0], shape[1])
            return xt if (xt.dtype == torch.float32 and xt.is_contiguous()) else xt.float().contiguous()
        return torch.from_numpy(np.ascontiguousarray(x[:, :, 0], dtype=np.float32))

    def predict(self, x, batch_size=32, verbose=0):
        """model.predict(x): eval-mode forward (moving BN statistics, no dropout).  x: numpy (N, L, 1) of any float
        dtype (cast to float32 like Keras' floatx) or a torch CPU tensor.  Returns numpy float32.  ``batch_size``
        is accepted for compatibility; results do not depend on it."""
        import torch
        if isinstance(x, torch.Tensor):
            xt = self._host_batch(x)
            count = xt.shape[0]
        else:
            x = np.asarray(x)
            self._check_input_shape(x.shape)
            count = x.shape[0]
        if count == 0:           # an empty batch yields an empty result (as numpy-style APIs do); nothing to launch
            units = self._head["units"] if self._head is not None else self.embedding_dimension
            return np.zeros((0, units), dtype=np.float32)
        eng = self._get_engine()
        if isinstance(x, torch.Tensor):
            emb = self._embed_pipelined(xt, eng)
        else:
            emb = torch.empty((count, self.embedding_dimension), dtype=torch.float32, device=eng.device)
            self._stage_numpy(x[:, :, 0], eng,
                              lambda xin, lo, hi, base: eng.forward(xin[lo:hi], out=emb[base + lo:base + hi]))
        if self._head is not None:
            emb = self._apply_head(emb)
        return emb.cpu().numpy()

    def _embed_pipelined(self, xt, eng):
        """Host batch (N, L) -> device embeddings.  A pinned batch is cut into a small first chunk and the rest
        (``_pipeline_plan``): the host->device copies run back to back on a side stream while the main stream
        embeds every chunk as soon as it has landed."""
        import torch
        n, length = xt.shape
        dev = eng.device
        out = torch.empty((n, self.embedding_dimension), dtype=torch.float32, device=dev)
        rows = max(1, min(n, _PIPELINE_BUFFER_BYTES // (4 * length)))   # device input buffer, grow-only
        if not xt.is_pinned() and n <= _PREDICT_CHUNK:
            eng.forward(xt.to(dev, non_blocking=True), out=out)
            return out
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._copy_buf = None
        if self._copy_buf is None or self._copy_buf.numel() < rows * length:
            self._copy_buf = torch.empty(rows * length, dtype=torch.float32, device=dev)
        xin = self._copy_buf[:rows * length].view(rows, length)
        main = torch.cuda.current_stream(dev)
        for base in range(0, n, rows):
            m = min(rows, n - base)
            self._copy_stream.wait_stream(main)     # earlier kernels that read the buffer are done
            for lo, hi in _pipeline_plan(m, xt.is_pinned()):
                with torch.cuda.stream(self._copy_stream):
                    xin[lo:hi].copy_(xt[base + lo:base + hi], non_blocking=True)
                    copied = torch.cuda.Event()
                    copied.record(self._copy_stream)
                main.wait_event(copied)
                eng.forward(xin[lo:hi], out=out[base + lo:base + hi])
        return out

    def _device_input(self, eng, rows, length):
        import torch
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=eng.device)
            self._copy_buf = None
        if self._copy_buf is None or self._copy_buf.numel() < rows * length:
            self._copy_buf = torch.empty(rows * length, dtype=torch.float32, device=eng.device)
        return self._copy_buf[:rows * length].view(rows, length)

    def _stage_numpy(self, src, eng, on_chunk=None, xin=None):
        """numpy (n, length), any dtype -> float32 on the device, through ``_HostStage``: cast on worker threads into
        pinned chunks, asynchronous copies on the side stream.  ``on_chunk(xin, lo, hi, base)`` is called once rows
        [lo, hi) of the device buffer ``xin`` -- batch rows [base + lo, base + hi) -- are ordered before the main
        stream; a batch larger than the staging buffer goes through it in passes (``base`` > 0).  With a caller-owned
        ``xin`` (n rows) and no callback the main stream just waits for the whole batch."""
        import torch
        n, length = src.shape
        if getattr(self, "_host_stage", None) is None:
            self._host_stage = _HostStage()
        if xin is None:
            rows = max(1, min(n, _PIPELINE_BUFFER_BYTES // (4 * length)))
            xin = self._device_input(eng, rows, length)
        else:
            rows = n
            self._device_input(eng, 1, 1)               # makes sure the copy stream exists
        main = torch.cuda.current_stream(eng.device)
        for base in range(0, n, rows):
            m = min(rows, n - base)
            self._copy_stream.wait_stream(main)         # earlier kernels that read the buffer are done
            for lo, hi, view, slot in self._host_stage.chunks(src[base:base + m]):
                with torch.cuda.stream(self._copy_stream):
                    xin[lo:hi].copy_(view, non_blocking=True)
                    copied = torch.cuda.Event()
                    copied.record(self._copy_stream)
                slot[1] = copied
                if on_chunk is not None:
                    main.wait_event(copied)
                    on_chunk(xin, lo, hi, base)
            if on_chunk is None:
                main.wait_stream(self._copy_stream)
        return xin

    def predict_raw(self, x, downsampling=4, whitening=True):
        """Embeddings of RAW clips (N, T, 1) (e.g. 48000 samples of 16 kHz audio): equivalent to
        ``predict(preprocess_instances(downsampling, whitening)(x))`` with the decimation and whitening of
        voicemap/utils.py:22-34 done on the device inside block 1 (x is one whiten() batch)."""
        import torch
        shape_check, self.input_shape = self.input_shape, None   # the raw length differs from the model's input
        try:
            xt = self._host_batch(x)
        finally:
            self.input_shape = shape_check
        eng = self._get_engine()
        emb = eng.forward_raw(xt.to(eng.device, non_blocking=True), downsampling, 1 if whitening else 0)
        if self._head is not None:
            emb = self._apply_head(emb)
        return emb.cpu().numpy()

    def _apply_head(self, emb):
        # classifier head Dense(num_classes, softmax): adjacent to the hot path, device-side
        import torch
        w = torch.from_numpy(self.weights["head_kernel"]).to(emb.device)
        b = torch.from_numpy(self.weights["head_bias"]).to(emb.device)
        z = emb @ w + b
        act = self._head["activation"]
        if act == "softmax":
            return torch.softmax(z, dim=-1)
        if act == "sigmoid":
            return torch.sigmoid(z)
        return z


class SiameseModel(_ModelBase):
    """Model returned by build_siamese_net (voicemap/models.py:44-81): the shared encoder on two inputs, a distance
    layer and Dense(1, sigmoid).  ``layers[2]`` is the encoder (voicemap/utils.py:141)."""

    def __init__(self, encoder, input_shape, distance_metric, seed=None):
        self.name = "model_1"
        self.encoder = encoder
        encoder._siamese_owner = self        # host-side weight changes of the encoder invalidate this model's trainer
        self.input_shape = tuple(input_shape)
        self.distance_metric = distance_metric
        rng = np.random.default_rng(seed)
        emb = encoder.embedding_dimension
        in_dim = emb if distance_metric == 'weighted_l1' else 1
        self.head_weights = OrderedDict(
            head_kernel=_glorot_uniform((in_dim, 1), rng), head_bias=np.zeros((1,), np.float32))
        if distance_metric == 'weighted_l1':
            mid = [LayerInfo("subtract_1", "Subtract", {}, (), self), LayerInfo("lambda_1", "Lambda", {}, (), self)]
        else:
            mid = [LayerInfo("subtract_embeddings", "Subtract", {}, (), self),
                   LayerInfo("euclidean_distance", "Lambda", {}, (), self)]
        self.layers = [LayerInfo("input_1", "InputLayer", {}, (), self), LayerInfo("input_2", "InputLayer", {}, (), self),
                       encoder, *mid,
                       LayerInfo("dense_2", "Dense", dict(units=1, activation="sigmoid"),
                                 ("head_kernel", "head_bias"), self)]
        self._head_dev = None
        # The head works on DIFFERENCES of embeddings: e1 - e2 cancels most of their magnitude, which amplifies the
        # encoder's rounding error by |e| / |e1 - e2| before the sigmoid.  Pair probabilities and the losses computed
        # from them are held to 1e-4 of the reference, which the fp16 x 3 arithmetic meets and the
        # fp16 + fp8 mode (the encoder's default for embeddings) does not always: the siamese eval path runs precision 3.
        self.precision = 3

    def _weight(self, name):
        return self.head_weights[name]

    def count_params(self):
        return self.encoder.count_params() + int(sum(w.size for w in self.head_weights.values()))

    def get_weights(self):
        return self.encoder.get_weights() + [w.copy() for w in self.head_weights.values()]

    def set_weights(self, weights):
        weights = list(weights)
        n_enc = len(self.encoder.weights)
        if len(weights) != n_enc + 2:
            raise ValueError(f"expected {n_enc + 2} weight arrays, got {len(weights)}")
        self.encoder.set_weights(weights[:n_enc])
        for (name, old), new in zip(self.head_weights.items(), weights[n_enc:]):
            new = np.asarray(new, np.float32)
            if new.shape != old.shape:
                raise ValueError(f"{name}: expected {old.shape}, got {new.shape}")
            self.head_weights[name] = new.copy()
        self._head_dev = None
        self._trainer = None

    def get_config(self):
        return dict(kind="siamese", encoder=self.encoder.get_config(), input_shape=self.input_shape,
                    distance_metric=self.distance_metric)

    def _keras_config(self):
        shape = [None] + [int(v) for v in self.input_shape]
        inp = lambda n: {"class_name": "InputLayer", "name": n, "inbound_nodes": [], "config": {
            "name": n, "dtype": "float32", "sparse": False, "batch_input_shape": shape}}
        # the distance layers are Python lambdas (voicemap/models.py:57,66); Keras stores them as interpreter-specific
        # bytecode, which cannot be produced here: ``function`` is null, so the file round-trips through this
        # package's load_model and keras' load_weights, but not through keras' load_model
        lam = {"class_name": "Lambda", "name": "lambda_1", "inbound_nodes": [[["subtract_1", 0, 0, {}]]], "config": {
            "name": "lambda_1", "trainable": True, "function": None, "function_type": "lambda", "arguments": {},
            "output_shape": None, "output_shape_type": "raw", "voicemap_distance_metric": self.distance_metric}}
        head = _dense_config("dense_2", 1, "sigmoid")
        head.update(name="dense_2", inbound_nodes=[[["lambda_1", 0, 0, {}]]])
        layers = [inp("input_1"), inp("input_2"),
                  {"class_name": "Sequential", "name": "sequential_1", "config": self.encoder._keras_layer_configs(),
                   "inbound_nodes": [[["input_1", 0, 0, {}]], [["input_2", 0, 0, {}]]]},
                  {"class_name": "Subtract", "name": "subtract_1", "config": {"name": "subtract_1", "trainable": True},
                   "inbound_nodes": [[["sequential_1", 1, 0, {}], ["sequential_1", 2, 0, {}]]]},
                  lam, head]
        return {"class_name": "Model", "config": {
            "name": "model_1", "layers": layers, "input_layers": [["input_1", 0, 0], ["input_2", 0, 0]],
            "output_layers": [["dense_2", 0, 0]]}}

    def _keras_layers(self):
        enc = [(f"sequential_1/{w}", a) for _, w, a in self.encoder._keras_weight_items(trainable_first=True)]
        head = [("dense_2/kernel:0", self.head_weights["head_kernel"]), ("dense_2/bias:0", self.head_weights["head_bias"])]
        return [("input_1", []), ("input_2", []), ("sequential_1", enc), ("subtract_1", []), ("lambda_1", []),
                ("dense_2", head)]

    def _clone(self):
        return SiameseModel(self.encoder._clone(), self.input_shape, self.distance_metric)

    def _head_device(self, device):
        import torch
        if self._head_dev is None or self._head_dev[0].device != device:
            self._head_dev = (torch.from_numpy(self.head_weights["head_kernel"].reshape(-1).copy()).to(device),
                              torch.from_numpy(self.head_weights["head_bias"].copy()).to(device))
        return self._head_dev

    def predict(self, x, batch_size=32, verbose=0):
        """siamese.predict([input_1, input_2]) -> (N, 1) sigmoid outputs (voicemap/utils.py:133)."""
        import torch
        from .engine import pair_head_loss
        if not isinstance(x, (list, tuple)) or len(x) != 2:
            raise ValueError("Error when checking model input: the siamese network expects a list of 2 arrays")
        sides = []
        for side in x:          # torch CPU tensors go through _host_batch; numpy batches are cast while they are staged
            if isinstance(side, torch.Tensor):
                sides.append(self.encoder._host_batch(side))
            else:
                side = np.asarray(side)
                self.encoder._check_input_shape(side.shape)
                sides.append(side[:, :, 0])
        x1, x2 = sides
        if tuple(x1.shape) != tuple(x2.shape):
            raise ValueError(f"siamese inputs must have the same shape, got {tuple(x1.shape)} and {tuple(x2.shape)}")
        if (x1.shape[1], 1) != self.input_shape:
            raise ValueError(f"Error when checking input: expected input_1 to have shape {self.input_shape} but got "
                             f"array with shape {(x1.shape[1], 1)}")
        if x1.shape[0] == 0:
            return np.zeros((0, 1), dtype=np.float32)
        eng = self.encoder._get_engine()
        w, b = self._head_device(eng.device)
        outs = []
        half = _PREDICT_CHUNK // 2
        for i in range(0, x1.shape[0], half):
            n = min(half, x1.shape[0] - i)
            # both branches share weights and eval-mode BN: run them as one 2n-clip batch
            xb = torch.empty((2 * n, x1.shape[1]), dtype=torch.float32, device=eng.device)
            for part, dst in ((x1, xb[:n]), (x2, xb[n:])):
                if isinstance(part, torch.Tensor):
                    dst.copy_(part[i:i + n], non_blocking=True)
                else:
                    self.encoder._stage_numpy(part[i:i + n], eng, xin=dst)
            emb = eng.forward(xb, precision=self.precision)
            prob, _, _ = pair_head_loss(emb[:n], emb[n:], w, b, self.distance_metric)
            outs.append(prob)
        out = outs[0] if len(outs) == 1 else torch.cat(outs, dim=0)
        return out.cpu().numpy()

    def test_on_batch(self, x, y, sample_weight=None):
        """keras ``Model.test_on_batch``: eval-mode loss of the compiled model on one batch of pairs (and accuracy when
        compiled with ``metrics=['accuracy']``).  Encoder, head and loss run on the device: both branches as one
        2N-clip launch, then ONE launch of the fused head + loss kernel (``vm_pair_head_loss_fwd``), which is the
        arithmetic the training step uses.  Reference: voicemap/models.py:64-69, voicemap/utils.py:77-85, keras
        binary_crossentropy (experiments/train_siamese.py:57)."""
        import torch
        from .engine import pair_head_loss
        if sample_weight is not None:
            raise NotImplementedError("sample_weight is not used by voicemap's scripts")
        if self.loss is None:
            raise RuntimeError("You must compile a model before training/testing. Use `model.compile(optimizer, loss)`.")
        loss = {"contrastive_loss": "contrastive", "binary_crossentropy": "binary_crossentropy"}.get(self.loss)
        if loss is None:
            raise NotImplementedError(f"siamese evaluation with loss {self.loss!r}")
        if not isinstance(x, (list, tuple)) or len(x) != 2:
            raise ValueError("Error when checking model input: the siamese network expects a list of 2 arrays")
        sides = []
        for side in x:
            side = np.asarray(side)
            self.encoder._check_input_shape(side.shape)
            sides.append(side[:, :, 0])
        n, length = sides[0].shape
        if sides[1].shape != (n, length):
            raise ValueError("siamese inputs must have the same shape")
        labels = np.asarray(y, dtype=np.float32).reshape(-1)
        if labels.shape[0] != n:
            raise ValueError(f"expected {n} labels, got {labels.shape[0]}")
        eng = self.encoder._get_engine()
        w, b = self._head_device(eng.device)
        xb = torch.empty((2 * n, length), dtype=torch.float32, device=eng.device)
        for part, dst in zip(sides, (xb[:n], xb[n:])):
            self.encoder._stage_numpy(part, eng, xin=dst)
        yt = torch.from_numpy(labels).to(eng.device, non_blocking=True)
        emb = eng.forward(xb, precision=self.precision)
        prob, _, lossv = pair_head_loss(emb[:n], emb[n:], w, b, self.distance_metric, y_true=yt, loss=loss)
        out = [float(lossv.item())]
        if any(m in ("accuracy", "acc") for m in (self.metrics or [])):
            out.append(float(((prob.reshape(-1) > 0.5).to(torch.float32) == yt).to(torch.float32).mean().item()))
        return out[0] if len(out) == 1 else out


# --------------------------------------------------------------------------------------------------------
# the reference's public builders
# --------------------------------------------------------------------------------------------------------
def get_baseline_convolutional_encoder(filters, embedding_dimension, input_shape=None, dropout=0.05, first_pool=4):
    """voicemap/models.py:6-41: 4 x [Conv1D+ReLU -> BatchNorm -> SpatialDropout1D -> MaxPool] ->
    GlobalMaxPool1D -> Dense(embedding_dimension).  ``first_pool`` (not in the reference signature; default = the
    reference's 4) selects the older architecture of the shipped checkpoints when set to 2."""
    return EncoderModel(filters, embedding_dimension, input_shape=input_shape, dropout=dropout, first_pool=first_pool)


def build_siamese_net(encoder, input_shape, distance_metric='uniform_euclidean'):
    """voicemap/models.py:44-81."""
    assert distance_metric in DISTANCE_METRICS
    if distance_metric not in ('weighted_l1', 'uniform_euclidean'):
        raise NotImplementedError  # voicemap/models.py:70-77
    return SiameseModel(encoder, input_shape, distance_metric)


def _load_keras_hdf5(filepath):
    """Keras 2.2.x full-model HDF5 checkpoint (what the reference's ModelCheckpoint writes,
    experiments/train_siamese.py:81-87) -> EncoderModel / SiameseModel.  Weights are matched by name, so both the
    nested (siamese: trainable first, moving statistics last) and the flat Sequential orders load."""
    import warnings
    from .keras_hdf5 import load_keras_weights
    cfg, layers = load_keras_weights(filepath)
    flat = OrderedDict()
    for _, ws in layers:
        for wname, arr in ws:
            flat[wname.split(":")[0]] = np.asarray(arr, dtype=np.float32)

    def find(suffix):
        hits = [k for k in flat if k.endswith(suffix)]
        if len(hits) != 1:
            raise ValueError(f"checkpoint has {len(hits)} weights matching {suffix!r}")
        return flat[hits[0]]

    filters = find("conv1d_1/kernel").shape[2]
    emb = find("dense_1/kernel").shape[1]
    named = OrderedDict()
    for i in range(1, 5):
        named[f"conv{i}_kernel"] = find(f"conv1d_{i}/kernel")
        named[f"conv{i}_bias"] = find(f"conv1d_{i}/bias")
        named[f"bn{i}_gamma"] = find(f"batch_normalization_{i}/gamma")
        named[f"bn{i}_beta"] = find(f"batch_normalization_{i}/beta")
        named[f"bn{i}_mean"] = find(f"batch_normalization_{i}/moving_mean")
        named[f"bn{i}_var"] = find(f"batch_normalization_{i}/moving_variance")
    named["dense_kernel"], named["dense_bias"] = find("dense_1/kernel"), find("dense_1/bias")
    pools = []

    def walk(o):
        if isinstance(o, dict):
            if o.get("class_name") == "MaxPooling1D":
                pools.append(tuple(o["config"].get("pool_size", ())))
            for v in o.values():
                walk(v)
        elif isinstance(o, list):
            for v in o:
                walk(v)

    walk(cfg)
    first_pool = int(pools[0][0]) if pools and pools[0] else 4
    if first_pool not in (2, 4) or any(p != (2,) for p in pools[1:4]):
        raise ValueError(f"checkpoint pooling sizes {pools} are not a voicemap encoder (4,2,2,2 or 2,2,2,2)")
    is_siamese = cfg.get("class_name") == "Model"
    drops = []

    def walk_drop(o):
        if isinstance(o, dict):
            if o.get("class_name") == "SpatialDropout1D":
                drops.append(float(o["config"].get("rate", 0.05)))
            for v in o.values():
                walk_drop(v)
        elif isinstance(o, list):
            for v in o:
                walk_drop(v)

    walk_drop(cfg)
    lshape = None

    def walk_shape(o):
        nonlocal lshape
        if isinstance(o, dict):
            if lshape is None and o.get("batch_input_shape"):
                lshape = tuple(o["batch_input_shape"][1:])
            for v in o.values():
                walk_shape(v)
        elif isinstance(o, list):
            for v in o:
                walk_shape(v)

    walk_shape(cfg)
    lshape = lshape if lshape and lshape[0] else None
    # older checkpoints (models/n_seconds/*.hdf5) were trained with a first MaxPool1D of 2
    enc = EncoderModel(filters, emb, input_shape=None if is_siamese else lshape,
                       dropout=drops[0] if drops else 0.05, first_pool=first_pool)
    enc.set_named_weights(named)
    if not is_siamese:
        if any(k.endswith("dense_2/kernel") for k in flat):     # classifier: Dense(num_classes, softmax) on top
            act = "softmax"
            seq = cfg.get("config")
            for layer in (seq if isinstance(seq, list) else seq.get("layers", [])):
                if layer.get("config", {}).get("name") == "dense_2":
                    act = layer["config"].get("activation", act)
            enc.add(Dense(find("dense_2/kernel").shape[1], activation=act))
            enc.set_named_weights({"head_kernel": find("dense_2/kernel"), "head_bias": find("dense_2/bias")})
        return enc
    head_k = find("dense_2/kernel")
    stored = []

    def walk_metric(o):
        if isinstance(o, dict):
            if "voicemap_distance_metric" in o:
                stored.append(o["voicemap_distance_metric"])
            for v in o.values():
                walk_metric(v)
        elif isinstance(o, list):
            for v in o:
                walk_metric(v)

    walk_metric(cfg)
    if stored and stored[0] in ("weighted_l1", "uniform_euclidean"):
        metric = stored[0]               # written by this package's save(); Keras' own files carry only the lambda
    else:                                # a Keras checkpoint: the head kernel's shape tells the two apart (emb > 1)
        metric = "weighted_l1" if head_k.shape[0] == emb and emb != 1 else "uniform_euclidean"
    m = SiameseModel(enc, lshape if lshape else (12000, 1), metric)
    m.head_weights["head_kernel"] = head_k.reshape(m.head_weights["head_kernel"].shape).copy()
    m.head_weights["head_bias"] = find("dense_2/bias").copy()
    return m


def _compile_from_training_config(model, filepath):
    """keras.models.load_model returns a compiled model when the file carries a training_config: same here, for the
    losses / optimizer this package implements (anything else leaves the model un-compiled, with a warning)."""
    import warnings
    from .keras_hdf5 import load_training_config
    tc = load_training_config(filepath)
    if not tc:
        return model
    loss = tc.get("loss")
    opt_cfg = (tc.get("optimizer_config") or {})
    if opt_cfg.get("class_name") != "Adam" or loss not in ("binary_crossentropy", "categorical_crossentropy",
                                                         "contrastive_loss"):
        warnings.warn(f"checkpoint training_config (loss {loss!r}, optimizer {opt_cfg.get('class_name')!r}) is not "
                      f"restored: compile the model before training")
        return model
    cfg = dict(opt_cfg.get("config") or {})
    kwargs = {k: cfg[k] for k in ("lr", "beta_1", "beta_2", "epsilon", "decay", "clipnorm") if cfg.get(k) is not None}
    model.compile(loss=loss, optimizer=Adam(**kwargs), metrics=list(tc.get("metrics") or []))
    return model


def load_model(filepath, custom_objects=None):
    """keras.models.load_model call site: experiments/k_way_accuracy.py:45.  Reads this package's npz container
    (counterpart of ``save``) and Keras 2.2.x HDF5 checkpoints (pure-Python reader, no h5py)."""
    with open(filepath, "rb") as fh:
        magic = fh.read(8)
    if magic == b"\x89HDF\r\n\x1a\n":
        return _compile_from_training_config(_load_keras_hdf5(filepath), filepath)
    with np.load(filepath) as z:
        cfg = json.loads(bytes(z["config"]).decode())
        weights = [z[f"w{i}"] for i in range(len(z.files) - 1)]
    if cfg["kind"] == "encoder":
        m = EncoderModel(cfg["filters"], cfg["embedding_dimension"], cfg["input_shape"], cfg["dropout"],
                         first_pool=cfg.get("first_pool", 4))
        if cfg.get("head"):
            m.add(Dense(cfg["head"]["units"], activation=cfg["head"]["activation"]))
    else:
        e = cfg["encoder"]
        enc = EncoderModel(e["filters"], e["embedding_dimension"], e["input_shape"], e["dropout"],
                           first_pool=e.get("first_pool", 4))
        m = SiameseModel(enc, cfg["input_shape"], cfg["distance_metric"])
    m.set_weights(weights)
    return m
