"""A numpy stand-in for the Keras symbols that voicemap/models.py and voicemap/utils.py touch, so that the REFERENCE'S
OWN SOURCE FILES can be imported and executed in this image (which has neither Keras nor TensorFlow) to generate golden
vectors -- see make_reference_golden.py.  Test infrastructure: nothing in the product imports it, and it is not the
oracle either (the oracle is oracle/voicemap_oracle.py, torch; this file shares no code with it).

What executing the reference through this file pins, and what it does not:
  * pinned by the reference's code itself: the architecture (layer order, filter multipliers, kernel sizes, padding
    mode, pool sizes, embedding width), the siamese wiring and both distance heads incl. the K.* lambdas, the
    contrastive loss expression, whitening / decimation, the n-shot decision rules, `layers[2]`, clone/pop;
  * NOT pinned: the arithmetic inside each Keras layer.  It is restated here from Keras 2.2.2 / TF 1.10 semantics
    (SURVEY.md 8(c)): cross-correlation Conv1D with TF 'SAME' padding (K-1 zeros, floor half on the left), BatchNorm in
    inference mode with epsilon 1e-3 on the last axis, dropout = identity outside training, VALID max pooling with
    stride = pool size by default, K.sqrt clipping its argument at 0.

Everything computes in float64 (or `set_floatx('float32')`), eagerly, on numpy arrays; a tiny symbolic layer (`Input`
nodes) exists only so that `build_siamese_net`'s functional-API calls and `Model(inputs, outputs)` work.
"""
import sys
import types

import numpy as np

_FLOATX = [np.float64]


def set_floatx(name):
    _FLOATX[0] = np.dtype(name).type


def _f(x):
    return np.asarray(x, dtype=_FLOATX[0])


# ----------------------------------------------------------------------------------------------- symbolic glue
class Node:
    """Output of `layer(parents)` in a functional-API graph; `layer is None` marks an Input."""

    def __init__(self, layer, parents, shape=None):
        self.layer, self.parents, self.shape = layer, parents, shape

    def evaluate(self, feeds):
        if id(self) in feeds:
            return feeds[id(self)]
        args = [p.evaluate(feeds) for p in self.parents]
        value = self.layer.compute(args if self.layer.takes_list else args[0])
        feeds[id(self)] = value
        return value


def _is_symbolic(x):
    return isinstance(x, Node) or (isinstance(x, (list, tuple)) and len(x) > 0 and isinstance(x[0], Node))


class Layer:
    takes_list = False
    n_weights = 0
    _counters = {}

    def __init__(self, name=None, **_):
        kind = type(self).__name__.lower()
        Layer._counters[kind] = Layer._counters.get(kind, 0) + 1
        self.name = name or '{}_{}'.format(kind, Layer._counters[kind])
        self.weights = [None] * self.n_weights

    def __call__(self, x):
        if _is_symbolic(x):
            return Node(self, list(x) if isinstance(x, (list, tuple)) else [x])
        return self.compute(x)

    def get_weights(self):
        return list(self.weights)

    def set_weights(self, weights):
        assert len(weights) == self.n_weights, (self.name, len(weights), self.n_weights)
        self.weights = [_f(w) for w in weights]

    def get_config(self):
        return {}

    def clone(self):
        return type(self)(**self.get_config())


def Input(shape=None, **_):
    return Node(None, [], shape)


# ----------------------------------------------------------------------------------------------- layers
def _activation(name, z):
    if name is None or name == 'linear':
        return z
    if name == 'relu':
        return np.maximum(z, 0)
    if name == 'sigmoid':
        return 1.0 / (1.0 + np.exp(-z))
    if name == 'softmax':
        e = np.exp(z - z.max(axis=-1, keepdims=True))
        return e / e.sum(axis=-1, keepdims=True)
    raise NotImplementedError(name)


class Conv1D(Layer):
    n_weights = 2  # kernel (K, Cin, Cout), bias (Cout,)

    def __init__(self, filters, kernel_size, strides=1, padding='valid', activation=None, input_shape=None, **kw):
        super().__init__(**kw)
        assert strides == 1
        self.filters, self.kernel_size, self.padding, self.activation = filters, int(kernel_size), padding, activation
        self.input_shape = input_shape

    def get_config(self):
        return dict(filters=self.filters, kernel_size=self.kernel_size, padding=self.padding,
                    activation=self.activation, input_shape=self.input_shape)

    def compute(self, x):
        kernel, bias = self.weights
        x = _f(x)
        k = self.kernel_size
        assert kernel.shape[0] == k and kernel.shape[1] == x.shape[2] and kernel.shape[2] == self.filters
        if self.padding == 'same':       # TF 'SAME', stride 1: K - 1 zeros in total, the smaller half in front
            left = (k - 1) // 2
            x = np.pad(x, ((0, 0), (left, k - 1 - left), (0, 0)))
        else:
            assert self.padding == 'valid'
        out_len = x.shape[1] - k + 1
        z = np.zeros((x.shape[0], out_len, self.filters), dtype=x.dtype)
        for tap in range(k):             # cross-correlation: no kernel flip
            z += x[:, tap:tap + out_len, :] @ kernel[tap]
        return _activation(self.activation, z + bias)


class BatchNormalization(Layer):
    n_weights = 4  # gamma, beta, moving_mean, moving_variance (Keras' get_weights order)

    def __init__(self, axis=-1, momentum=0.99, epsilon=1e-3, **kw):
        super().__init__(**kw)
        assert axis == -1
        self.momentum, self.epsilon = momentum, epsilon

    def get_config(self):
        return dict(momentum=self.momentum, epsilon=self.epsilon)

    def compute(self, x):                # inference phase (`predict`): moving statistics
        gamma, beta, mean, var = self.weights
        inv = gamma / np.sqrt(var + _f(self.epsilon))
        return _f(x) * inv + (beta - mean * inv)


class SpatialDropout1D(Layer):
    def __init__(self, rate, **kw):
        super().__init__(**kw)
        self.rate = rate

    def get_config(self):
        return dict(rate=self.rate)

    def compute(self, x):                # dropout is the identity outside the training phase
        return x


class MaxPool1D(Layer):
    def __init__(self, pool_size=2, strides=None, padding='valid', **kw):
        super().__init__(**kw)
        assert padding == 'valid'
        self.pool_size = int(pool_size)
        self.strides = int(strides) if strides is not None else self.pool_size

    def get_config(self):
        return dict(pool_size=self.pool_size, strides=self.strides)

    def compute(self, x):
        p, s = self.pool_size, self.strides
        out_len = (x.shape[1] - p) // s + 1
        windows = np.stack([x[:, j:j + (out_len - 1) * s + 1:s, :] for j in range(p)], axis=0)
        return windows.max(axis=0)


MaxPooling1D = MaxPool1D


class GlobalMaxPool1D(Layer):
    def compute(self, x):
        return x.max(axis=1)


GlobalMaxPooling1D = GlobalMaxPool1D


class Dense(Layer):
    n_weights = 2

    def __init__(self, units, activation=None, **kw):
        super().__init__(**kw)
        self.units, self.activation = units, activation

    def get_config(self):
        return dict(units=self.units, activation=self.activation)

    def compute(self, x):
        kernel, bias = self.weights
        return _activation(self.activation, _f(x) @ kernel + bias)


class Subtract(Layer):
    takes_list = True

    def compute(self, xs):
        return xs[0] - xs[1]


class Lambda(Layer):
    def __init__(self, function, **kw):
        super().__init__(**kw)
        self.function = function

    def compute(self, x):
        return self.function(x)


# ----------------------------------------------------------------------------------------------- models
class Sequential(Layer):
    def __init__(self, layers=None, **kw):
        super().__init__(**kw)
        self.layers = list(layers or [])

    def add(self, layer):
        self.layers.append(layer)

    def pop(self):
        self.layers.pop()

    def compute(self, x):
        for layer in self.layers:
            x = layer.compute(x)
        return x

    def predict(self, x, **_):
        return self.compute(_f(x))

    def get_weights(self):
        return [w for layer in self.layers for w in layer.get_weights()]

    def set_weights(self, weights):
        weights = list(weights)
        for layer in self.layers:
            layer.set_weights([weights.pop(0) for _ in range(layer.n_weights)])
        assert not weights


class Model:
    """Functional model.  `layers` follows Keras' ordering for this graph shape: the inputs, then every layer in
    depth order (for build_siamese_net: input_1, input_2, the shared encoder, subtract, lambda, dense -- the order the
    shipped checkpoint's `layer_names` records, which voicemap/utils.py:141 relies on with `model.layers[2]`)."""

    def __init__(self, inputs, outputs):
        self.inputs = list(inputs) if isinstance(inputs, (list, tuple)) else [inputs]
        self.output = outputs
        order, seen = [], set()

        def visit(node):
            for parent in node.parents:
                visit(parent)
            if node.layer is not None and id(node.layer) not in seen:
                seen.add(id(node.layer))
                order.append(node.layer)

        visit(outputs)
        self.layers = [Layer(name='input_{}'.format(i + 1)) for i in range(len(self.inputs))] + order

    def predict(self, xs, **_):
        xs = xs if isinstance(xs, (list, tuple)) else [xs]
        feeds = {id(node): _f(x) for node, x in zip(self.inputs, xs)}
        return self.output.evaluate(feeds)


def clone_model(model):
    """Same architecture, fresh (unset) weights -- callers follow with set_weights (voicemap/utils.py:143-145)."""
    assert isinstance(model, Sequential)
    return Sequential([layer.clone() for layer in model.layers])


# ----------------------------------------------------------------------------------------------- backend
class _Backend(types.ModuleType):
    @staticmethod
    def abs(x):
        return np.abs(x)

    @staticmethod
    def square(x):
        return np.square(x)

    @staticmethod
    def sqrt(x):                         # Keras' TF backend clips the argument to [0, inf) first
        return np.sqrt(np.clip(x, 0.0, np.inf))

    @staticmethod
    def sum(x, axis=None, keepdims=False):
        return np.sum(x, axis=axis, keepdims=keepdims)

    @staticmethod
    def mean(x, axis=None, keepdims=False):
        return np.mean(x, axis=axis, keepdims=keepdims)

    @staticmethod
    def maximum(x, y):
        return np.maximum(x, y)


class Callback:
    def __init__(self):
        self.model = None

    def set_model(self, model):
        self.model = model


def install():
    """Register the stand-in as `keras`, `keras.models`, `keras.layers`, `keras.backend`, `keras.callbacks`."""
    this = sys.modules[__name__]
    keras = types.ModuleType('keras')
    models = types.ModuleType('keras.models')
    models.Model, models.Sequential, models.clone_model = Model, Sequential, clone_model
    layers = types.ModuleType('keras.layers')
    for name in ('Conv1D', 'BatchNormalization', 'SpatialDropout1D', 'MaxPool1D', 'MaxPooling1D', 'GlobalMaxPool1D',
                 'GlobalMaxPooling1D', 'Dense', 'Subtract', 'Lambda', 'Input'):
        setattr(layers, name, getattr(this, name))
    backend = _Backend('keras.backend')
    callbacks = types.ModuleType('keras.callbacks')
    callbacks.Callback = Callback
    keras.models, keras.layers, keras.backend, keras.callbacks = models, layers, backend, callbacks
    for module in (keras, models, layers, backend, callbacks):
        sys.modules[module.__name__] = module
    return keras
