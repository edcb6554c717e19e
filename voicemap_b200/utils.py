"""``voicemap.utils`` for the B200 build -- the callers either side of the encoder: waveform
preprocessing, the contrastive loss, k-way n-shot evaluation and its per-epoch callback.  Names, arguments and
error behaviour follow voicemap/utils.py; the structure does not: preprocessing is an introspectable object (so
the engine can fuse decimation + whitening into block 1, ``EncoderModel.predict_raw``), and n-shot evaluation is
one pipeline (draw tasks -> embed -> score) that either calls ``predict`` per task like the reference or embeds
many tasks per launch.
"""
from __future__ import annotations

import numpy as np

from .keras_compat import Callback, clone_model

WHITENING_RMS = 0.038021


# ------------------------------------------------------------------------------------------------ preprocessing
def whiten(batch, rms=WHITENING_RMS):
    """Remove every clip's own mean, then multiply the whole batch by ONE scalar ``rms / sqrt(mean(batch ** 2))``
    taken over the un-centred batch (voicemap/utils.py:88-101; SURVEY.md F8 -- not a per-clip RMS)."""
    batch = np.asarray(batch)
    if batch.ndim != 3:
        raise ValueError('Input must be a 3D array of shape (n_segments, n_timesteps, 1).')
    gain = rms / np.sqrt(np.mean(np.square(batch)))
    return (batch - batch.mean(axis=1, keepdims=True)) * gain


class _InstancePreprocessor(object):
    """``instances[:, ::downsampling, :]`` then optionally ``whiten`` (voicemap/utils.py:22-34)."""

    def __init__(self, downsampling, whitening):
        self.downsampling = downsampling
        self.whitening = whitening

    def __call__(self, instances):
        out = instances[:, ::self.downsampling, :]
        return whiten(out) if self.whitening else out


def preprocess_instances(downsampling, whitening=True):
    """The project's canonical preprocessing (voicemap/utils.py:22-34): plain decimation by ``downsampling``
    followed by ``whiten``.  Returns a callable taking a (n, timesteps, 1) batch."""
    return _InstancePreprocessor(downsampling, whitening)


def _identity(x):
    return x


class BatchPreProcessor(object):
    """Apply one instance preprocessor (and optionally a target preprocessor) to classifier batches
    ``(inputs, targets)`` or siamese batches ``([input_1, input_2], targets)`` (voicemap/utils.py:37-74)."""

    def __init__(self, mode, instance_preprocessor, target_preprocessor=_identity):
        assert mode in ('siamese', 'classifier')
        self.mode = mode
        self.instance_preprocessor = instance_preprocessor
        self.target_preprocessor = target_preprocessor

    def __call__(self, batch):
        inputs, targets = batch
        if self.mode == 'siamese':
            inputs = [self.instance_preprocessor(side) for side in inputs]
        elif self.mode == 'classifier':
            inputs = self.instance_preprocessor(inputs)
        else:
            raise ValueError
        return inputs, self.target_preprocessor(targets)


# ------------------------------------------------------------------------------------------------ loss
def contrastive_loss(y_true, y_pred):
    """Hadsell et al. '06 with margin 1 (voicemap/utils.py:77-85): same-speaker pairs (y = 0) pay ``p ** 2``,
    different-speaker pairs (y = 1) pay ``max(1 - p, 0) ** 2``; mean over the batch.

    ``model.compile(loss=contrastive_loss)`` selects the fused CUDA head + loss kernel by identity of this function;
    a direct call evaluates the same expression with numpy."""
    y = np.asarray(y_true)
    p = np.asarray(y_pred)
    margin = 1
    pull = np.square(p)
    push = np.square(np.clip(margin - p, 0, None))
    return np.mean((1 - y) * pull + y * push)


# ------------------------------------------------------------------------------------------------ n-shot evaluation
def _per_class(values, k, n):
    """Mean over the n shots of each of the k classes; rows are ordered [class 1] * n + ... + [class k] * n."""
    return values.reshape(k, n, -1).mean(axis=1)


def _score_euclidean(query, support, k, n):
    return np.linalg.norm(_per_class(support, k, n) - query, axis=1)


def _score_cosine(query, support, k, n):
    # cosine distance between the query and the mean of each class's unit vectors (scipy cdist(..., 'cosine'))
    norms = np.linalg.norm(support, axis=1, keepdims=True)
    centre = _per_class(support / norms, k, n)
    return 1.0 - centre.dot(query) / (np.linalg.norm(centre, axis=1) * np.linalg.norm(query))


def _score_dot_product(query, support, k, n):
    # minus the projection of the query on (mean magnitude * mean direction) of each class
    norms = np.linalg.norm(support, axis=1, keepdims=True)
    centre = _per_class(support / norms, k, n) * _per_class(norms, k, n)
    return -centre.dot(query)


_SCORES = {'euclidean': _score_euclidean, 'cosine': _score_cosine, 'dot_product': _score_dot_product}


def _embedding_network(model, network_type):
    """The encoder inside ``model``: layer 2 of a siamese net, or a classifier without its softmax layer
    (voicemap/utils.py:140-147)."""
    if network_type == 'siamese':
        return model.layers[2]
    if network_type == 'classifier':
        stripped = clone_model(model)
        stripped.set_weights(model.get_weights())
        stripped.pop()
        return stripped
    raise ValueError('mode must be one of (siamese, classifier)')


def _task_inputs(dataset, preprocessor, k, n, pairwise):
    """Draw one task and preprocess it.  Pairwise (siamese, n = 1): the query repeated k times and the k support
    clips go through the batch preprocessor as one siamese batch, i.e. each side is whitened on its own
    (voicemap/utils.py:122-131).  Otherwise query (1 clip) and support (k*n clips) are preprocessed separately
    (:150-153)."""
    (query, _), (support, _) = dataset.build_n_shot_task(k, n)
    support = support[:, :, np.newaxis]
    if pairwise:
        repeated = np.repeat(query[np.newaxis, :, np.newaxis], k, axis=0)
        (left, right), _ = preprocessor(([repeated, support], []))
        return left, right
    instance = preprocessor.instance_preprocessor
    return instance(query.reshape(1, -1, 1)), instance(support)


def _solved_pairwise(model, tasks, k):
    """Tasks whose smallest siamese output belongs to support item 0, scoring all tasks of the list with one
    encoder launch and one head launch (one task: falls back to ``model.predict`` like the reference)."""
    if len(tasks) == 1:
        left, right = tasks[0]
        return int(np.argmin(model.predict([left, right])[:, 0]) == 0)
    from .engine import pair_head_loss
    t = len(tasks)
    encoder = model.encoder
    engine = encoder._get_engine()
    # eval-mode embeddings are independent of batch composition, so the query is embedded once per task
    clips = np.concatenate([left[:1] for left, _ in tasks] + [right for _, right in tasks], axis=0)
    emb = engine.forward(encoder._host_batch(clips).to(engine.device, non_blocking=True),
                         precision=getattr(model, 'precision', None))
    width = emb.shape[1]
    queries = emb[:t, None, :].expand(t, k, width).reshape(t * k, width).contiguous()
    kernel, bias = model._head_device(engine.device)
    prob, _, _ = pair_head_loss(queries, emb[t:].contiguous(), kernel, bias, model.distance_metric)
    return int((prob.reshape(t, k).argmin(dim=1) == 0).sum().item())


_DISTANCE_IDS = {'euclidean': 0, 'cosine': 1, 'dot_product': 2}   # VM_DISTANCE_* of include/voicemap_b200.h


def nshot_best_class_device(query_emb, support_emb, k, n, distance):
    """Device-side decision rule of voicemap/utils.py:156-212 for T tasks at once: ``query_emb`` CUDA (T, E),
    ``support_emb`` CUDA (T*k*n, E) ordered task by task, class by class.  Class means, the chosen distance and the
    arg-min all run in ``vm_nshot_score``; returns a CUDA int32 (T,) tensor of winning classes (0 = solved)."""
    import ctypes as C
    import torch
    from . import _lib
    lib = _lib.load()
    t, width = query_emb.shape
    if support_emb.shape != (t * k * n, width):
        raise ValueError('support embeddings must have shape (tasks * k * n, embedding_dimension)')
    query_emb, support_emb = query_emb.contiguous(), support_emb.contiguous()
    best = torch.empty((t,), dtype=torch.int32, device=query_emb.device)
    with torch.cuda.device(query_emb.device):
        rc = lib.vm_nshot_score(C.c_void_p(query_emb.data_ptr()), C.c_void_p(support_emb.data_ptr()), t, k, n, width,
                                _DISTANCE_IDS[distance], None, C.c_void_p(best.data_ptr()),
                                C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc, 'vm_nshot_score')
    return best


def _solved_by_embedding(encoder, tasks, k, n, distance):
    """Tasks whose nearest class (under ``distance``) is class 0.  Queries and supports of all tasks of the list are
    embedded by one encoder launch; class means, distances and the arg-min run on the device (``vm_nshot_score``) and
    only the count comes back.  (The numpy ``_SCORES`` above state the same rule on the host; tests hold the two
    together.)"""
    t = len(tasks)
    clips = np.concatenate([q for q, _ in tasks] + [s for _, s in tasks], axis=0)
    if not hasattr(encoder, '_get_engine'):
        # a foreign model object that only offers ``predict`` (duck-typed stand-ins in the host tests): its embeddings
        # are host arrays it computed itself; the rule is applied with the numpy statement above.  Models built by
        # this package never come here -- they have no host arithmetic to fall back to.
        emb = np.asarray(encoder.predict(clips))
        shots = emb[t:].reshape(t, k * n, -1)
        return sum(int(np.argmin(_SCORES[distance](emb[i], shots[i], k, n)) == 0) for i in range(t))
    import torch
    encoder._check_input_shape(clips.shape)
    engine = encoder._get_engine()
    emb = torch.empty((clips.shape[0], encoder.embedding_dimension), dtype=torch.float32, device=engine.device)
    encoder._stage_numpy(clips[:, :, 0], engine,
                         lambda xin, lo, hi, base: engine.forward(xin[lo:hi], out=emb[base + lo:base + hi]))
    best = nshot_best_class_device(emb[:t], emb[t:], k, n, distance)
    return int((best == 0).sum().item())


def n_shot_task_evaluation(model, dataset, preprocessor, num_tasks, n, k, network_type='siamese',
                           distance='euclidean', tasks_per_launch=1):
    """Number of correctly solved k-way n-shot tasks out of ``num_tasks`` (voicemap/utils.py:104-216).

    A task is solved when support item / class 0 -- the query's speaker by construction of
    ``dataset.build_n_shot_task`` -- is the closest: for a siamese network and n = 1 by the network's own pairwise
    output (smallest wins), otherwise by ``distance`` ('euclidean' | 'cosine' | 'dot_product') between the query
    embedding and the per-class mean embedding.

    ``tasks_per_launch`` > 1 (an extension) embeds that many tasks per encoder launch instead of issuing one or two
    tiny ``predict`` calls per task; tasks are drawn in the same order and scored by the same rule, so the count is
    the same for the same random state."""
    if n < 1:
        raise ValueError('n must be >= 1')
    pairwise = n == 1 and network_type == 'siamese'
    if pairwise:
        encoder, score = None, None
    else:
        encoder = _embedding_network(model, network_type)
        if distance not in _SCORES:
            raise ValueError('Distance must be in (euclidean, cosine, dot_product)')
        score = distance

    from tqdm import tqdm
    n_correct = 0
    progress = tqdm(total=num_tasks)
    for first in range(0, num_tasks, max(1, tasks_per_launch)):
        count = min(max(1, tasks_per_launch), num_tasks - first)
        tasks = [_task_inputs(dataset, preprocessor, k, n, pairwise) for _ in range(count)]
        if pairwise:
            n_correct += _solved_pairwise(model, tasks, k)
        else:
            n_correct += _solved_by_embedding(encoder, tasks, k, n, score)
        progress.update(count)
    progress.close()
    return n_correct


def n_shot_task_evaluation_batched(model, dataset, preprocessor, num_tasks, n, k, network_type='siamese',
                                   distance='euclidean', tasks_per_launch=64):
    """``n_shot_task_evaluation`` with many tasks per encoder launch (the reference spends its per-epoch evaluation in
    500 batch-5 forwards: experiments/train_siamese.py:75-78)."""
    return n_shot_task_evaluation(model, dataset, preprocessor, num_tasks, n, k, network_type=network_type,
                                  distance=distance, tasks_per_launch=tasks_per_launch)


class NShotEvaluationCallback(Callback):
    """After every epoch, run ``num_tasks`` k-way n-shot tasks on ``dataset`` and publish the accuracy as
    ``logs['val_{n}-shot_acc']`` for the callbacks that follow in the list (voicemap/utils.py:219-252).
    ``batch_tasks`` > 0 (an extension) evaluates that many tasks per encoder launch."""

    def __init__(self, num_tasks, n_shot, k_way, dataset, preprocessor=_identity, mode='siamese', batch_tasks=0):
        super(NShotEvaluationCallback, self).__init__()
        assert mode in ('siamese', 'classifier')
        self.num_tasks, self.n_shot, self.k_way = num_tasks, n_shot, k_way
        self.dataset = dataset
        self.preprocessor = preprocessor
        self.mode = mode
        self.batch_tasks = batch_tasks

    def on_epoch_end(self, epoch, logs=None):
        solved = n_shot_task_evaluation(self.model, self.dataset, self.preprocessor, self.num_tasks, self.n_shot,
                                        self.k_way, network_type=self.mode,
                                        tasks_per_launch=max(1, self.batch_tasks))
        key = 'val_{}-shot_acc'.format(self.n_shot)
        # data-parallel training: every rank has solved its own random tasks; the metric is their pooled accuracy, the
        # same number on every rank, so that callbacks steered by it (ReduceLROnPlateau, ModelCheckpoint) stay in step
        from .parallel import global_mean
        accuracy = global_mean(solved, self.num_tasks)
        if logs is not None:
            logs[key] = accuracy
        print('{}: {:.4f}'.format(key, accuracy))
