"""The slice of the Keras 2.2.2 object API that voicemap's scripts use on the models returned by
``voicemap.models``, re-implemented on top of the B200 engine.

Only what the reference calls is provided: ``Dense`` (as an argument to ``Sequential.add``), ``Adam``, the
callbacks ``CSVLogger`` / ``ModelCheckpoint`` / ``ReduceLROnPlateau`` / ``Callback``, ``Sequence``,
``to_categorical``, ``clone_model``, ``plot_model`` (no-op: graphviz is not part of the path).

Reference call sites: experiments/train_classifier.py:2-6,110-150; experiments/train_siamese.py:2-4,54-94;
voicemap/utils.py:2-3,142-145,219-252.
"""
from __future__ import annotations

import csv
import os

import numpy as np


# ---------------------------------------------------------------------------------------------- layers
class Dense:
    """keras.layers.Dense as used by ``classifier.add(Dense(num_classes, activation='softmax'))``
    (experiments/train_classifier.py:112)."""

    def __init__(self, units, activation=None, use_bias=True, name=None):
        if activation not in (None, "linear", "softmax", "sigmoid"):
            raise ValueError(f"activation {activation!r} is not used anywhere on the voicemap path")
        self.units = int(units)
        self.activation = activation
        self.use_bias = use_bias
        self.name = name


# ---------------------------------------------------------------------------------------------- optimizer
class Adam:
    """keras.optimizers.Adam(lr=1e-3, beta_1=.9, beta_2=.999, epsilon=None->1e-7, decay=0., clipnorm=...)
    (experiments/train_siamese.py:56, grid_search_siamese_network.py:56).  Holds hyper-parameters only; the
    update itself runs on the device."""

    def __init__(self, lr=0.001, beta_1=0.9, beta_2=0.999, epsilon=None, decay=0.0, amsgrad=False, clipnorm=None,
                 clipvalue=None):
        if amsgrad:
            raise NotImplementedError("amsgrad is not used by voicemap")
        if clipvalue is not None:
            raise NotImplementedError("clipvalue is not used by voicemap")
        self.lr = float(lr)
        self.beta_1 = float(beta_1)
        self.beta_2 = float(beta_2)
        self.epsilon = 1e-7 if epsilon is None else float(epsilon)
        self.decay = float(decay)
        self.clipnorm = None if clipnorm is None else float(clipnorm)
        self.iterations = 0

    def get_config(self):
        """keras optimizer_config of a saved model (training_config attribute)."""
        cfg = dict(lr=self.lr, beta_1=self.beta_1, beta_2=self.beta_2, decay=self.decay, epsilon=self.epsilon,
                   amsgrad=False)
        if self.clipnorm is not None:
            cfg["clipnorm"] = self.clipnorm
        return cfg


# ---------------------------------------------------------------------------------------------- utils
class Sequence:
    """keras.utils.Sequence protocol: __getitem__, __len__, on_epoch_end."""

    def __getitem__(self, index):
        raise NotImplementedError

    def __len__(self):
        raise NotImplementedError

    def on_epoch_end(self):
        pass

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]


def to_categorical(y, num_classes=None, dtype="float32"):
    """keras.utils.to_categorical (experiments/train_classifier.py:96)."""
    y = np.array(y, dtype="int")
    input_shape = y.shape
    if input_shape and input_shape[-1] == 1 and len(input_shape) > 1:
        input_shape = tuple(input_shape[:-1])
    y = y.ravel()
    if not num_classes:
        num_classes = int(np.max(y)) + 1
    n = y.shape[0]
    out = np.zeros((n, num_classes), dtype=dtype)
    out[np.arange(n), y] = 1
    return out.reshape(input_shape + (num_classes,))


def plot_model(model, to_file=None, show_shapes=False, **_):
    """Accepted and ignored (graphviz rendering is outside the path)."""
    return None


def clone_model(model):
    """keras.models.clone_model: same architecture, freshly initialised weights (voicemap/utils.py:143)."""
    return model._clone()


# ---------------------------------------------------------------------------------------------- callbacks
class Callback:
    def __init__(self):
        self.model = None
        self.params = {}

    def set_model(self, model):
        self.model = model

    def set_params(self, params):
        self.params = params

    def on_train_begin(self, logs=None):
        pass

    def on_train_end(self, logs=None):
        pass

    def on_epoch_begin(self, epoch, logs=None):
        pass

    def on_epoch_end(self, epoch, logs=None):
        pass

    def on_batch_begin(self, batch, logs=None):
        pass

    def on_batch_end(self, batch, logs=None):
        pass


class CSVLogger(Callback):
    """keras.callbacks.CSVLogger: one row per epoch, columns = sorted log keys of the first epoch."""

    def __init__(self, filename, separator=",", append=False):
        super().__init__()
        self.filename = filename
        self.sep = separator
        self.append = append
        self.keys = None
        self._file = None
        self._writer = None

    def on_train_begin(self, logs=None):
        os.makedirs(os.path.dirname(os.path.abspath(self.filename)), exist_ok=True)
        self._file = open(self.filename, "a" if self.append else "w", newline="")

    def on_epoch_end(self, epoch, logs=None):
        logs = logs or {}
        if self.keys is None:
            self.keys = sorted(logs.keys())
            self._writer = csv.DictWriter(self._file, fieldnames=["epoch"] + self.keys, delimiter=self.sep)
            self._writer.writeheader()
        row = {"epoch": epoch}
        row.update({k: logs.get(k, "NA") for k in self.keys})
        self._writer.writerow(row)
        self._file.flush()

    def on_train_end(self, logs=None):
        if self._file is not None:
            self._file.close()
            self._file = None


def _monitor_op(mode, monitor):
    if mode == "min" or (mode == "auto" and "acc" not in monitor):
        return (lambda a, b: a < b), np.inf
    return (lambda a, b: a > b), -np.inf


class ModelCheckpoint(Callback):
    """keras.callbacks.ModelCheckpoint(filepath, monitor, mode, save_best_only, verbose)
    (experiments/train_siamese.py:81-87).  Saves through ``model.save``: a Keras-layout HDF5 file for ``*.hdf5`` /
    ``*.h5`` paths (what the reference's scripts name their checkpoints), an npz container otherwise."""

    def __init__(self, filepath, monitor="val_loss", verbose=0, save_best_only=False, save_weights_only=False,
                 mode="auto", period=1):
        super().__init__()
        self.filepath = filepath
        self.monitor = monitor
        self.verbose = verbose
        self.save_best_only = save_best_only
        self.period = period
        self.epochs_since_last_save = 0
        self.monitor_op, self.best = _monitor_op(mode, monitor)

    def on_epoch_end(self, epoch, logs=None):
        logs = logs or {}
        self.epochs_since_last_save += 1
        if self.epochs_since_last_save < self.period:
            return
        self.epochs_since_last_save = 0
        filepath = self.filepath.format(epoch=epoch + 1, **logs)
        if self.save_best_only:
            current = logs.get(self.monitor)
            if current is None:
                print(f"Can save best model only with {self.monitor} available, skipping.")
                return
            if self.monitor_op(current, self.best):
                if self.verbose:
                    print(f"\nEpoch {epoch + 1:05d}: {self.monitor} improved from {self.best:0.5f} to "
                          f"{current:0.5f}, saving model to {filepath}")
                self.best = current
                self.model.save(filepath)
            elif self.verbose:
                print(f"\nEpoch {epoch + 1:05d}: {self.monitor} did not improve from {self.best:0.5f}")
        else:
            if self.verbose:
                print(f"\nEpoch {epoch + 1:05d}: saving model to {filepath}")
            self.model.save(filepath)


class ReduceLROnPlateau(Callback):
    """keras.callbacks.ReduceLROnPlateau defaults: factor .1, patience 10, min_delta 1e-4, cooldown 0, min_lr 0
    (experiments/train_siamese.py:88-92)."""

    def __init__(self, monitor="val_loss", factor=0.1, patience=10, verbose=0, mode="auto", min_delta=1e-4,
                 cooldown=0, min_lr=0):
        super().__init__()
        if factor >= 1.0:
            raise ValueError("ReduceLROnPlateau does not support a factor >= 1.0.")
        self.monitor, self.factor, self.patience, self.verbose = monitor, factor, patience, verbose
        self.min_delta, self.cooldown, self.min_lr = min_delta, cooldown, min_lr
        self.cooldown_counter = 0
        self.wait = 0
        if mode == "min" or (mode == "auto" and "acc" not in monitor):
            self.monitor_op = lambda a, b: a < b - self.min_delta
            self.best = np.inf
        else:
            self.monitor_op = lambda a, b: a > b + self.min_delta
            self.best = -np.inf

    def on_epoch_end(self, epoch, logs=None):
        logs = logs if logs is not None else {}
        opt = self.model.optimizer
        logs["lr"] = opt.lr
        current = logs.get(self.monitor)
        if current is None:
            return
        if self.cooldown_counter > 0:
            self.cooldown_counter -= 1
            self.wait = 0
        if self.monitor_op(current, self.best):
            self.best = current
            self.wait = 0
        elif self.cooldown_counter <= 0:
            self.wait += 1
            if self.wait >= self.patience:
                old_lr = float(opt.lr)
                if old_lr > self.min_lr:
                    opt.lr = max(old_lr * self.factor, self.min_lr)
                    if self.verbose:
                        print(f"\nEpoch {epoch + 1:05d}: ReduceLROnPlateau reducing learning rate to {opt.lr}.")
                    self.cooldown_counter = self.cooldown
                    self.wait = 0
