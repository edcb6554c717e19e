"""The handful of `K.*` calls of the reference's loss code (voicemap/utils.py:82-85), on numpy arrays."""
import numpy as np


def mean(x, axis=None, keepdims=False):
    return np.mean(x, axis=axis, keepdims=keepdims)


def square(x):
    return np.square(x)


def maximum(x, y):
    return np.maximum(x, y)


def abs(x):  # noqa: A001 - Keras' name
    return np.abs(x)


def sum(x, axis=None, keepdims=False):  # noqa: A001 - Keras' name
    return np.sum(x, axis=axis, keepdims=keepdims)


def sqrt(x):
    return np.sqrt(np.clip(x, 0.0, np.inf))
