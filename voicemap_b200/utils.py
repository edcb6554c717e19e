"""voicemap.utils for the B200 build: preprocessing, contrastive loss, n-shot evaluation and its callback,
with the reference's names and argument meaning (voicemap/utils.py).  Host-side numpy as in the reference;
the model calls inside go to the CUDA engine.
"""
from __future__ import annotations

import numpy as np
from tqdm import tqdm

from .keras_compat import Callback, clone_model


def preprocess_instances(downsampling, whitening=True):
    """This is the canonical preprocessing function for this project (voicemap/utils.py:22-34).

    1. Downsampling audio segments to desired sampling rate (plain decimation)
    2. Whiten audio segments to 0 mean and fixed RMS (aka volume)
    """
    def preprocess_instances_(instances):
        instances = instances[:, ::downsampling, :]
        if whitening:
            instances = whiten(instances)
        return instances

    return preprocess_instances_


class BatchPreProcessor(object):
    """Wrapper class for instance and label pre-processing (voicemap/utils.py:37-74).

    Pre-processes classifier-style batches (inputs, outputs) and siamese network-style batches
    ([input_1, input_2], outputs) identically.
    """
    def __init__(self, mode, instance_preprocessor, target_preprocessor=lambda x: x):
        assert mode in ('siamese', 'classifier')
        self.mode = mode
        self.instance_preprocessor = instance_preprocessor
        self.target_preprocessor = target_preprocessor

    def __call__(self, batch):
        if self.mode == 'siamese':
            ([input_1, input_2], labels) = batch
            input_1 = self.instance_preprocessor(input_1)
            input_2 = self.instance_preprocessor(input_2)
            labels = self.target_preprocessor(labels)
            return [input_1, input_2], labels
        elif self.mode == 'classifier':
            instances, labels = batch
            instances = self.instance_preprocessor(instances)
            labels = self.target_preprocessor(labels)
            return instances, labels
        else:
            raise ValueError


def contrastive_loss(y_true, y_pred):
    """Contrastive loss from Hadsell-et-al.'06 (voicemap/utils.py:77-85), margin 1, y: 0 = same speaker.

    Passing this function to ``model.compile(loss=contrastive_loss)`` selects the fused CUDA head+loss kernel;
    calling it directly evaluates the same expression on numpy arrays."""
    margin = 1
    y_true = np.asarray(y_true)
    y_pred = np.asarray(y_pred)
    return np.mean((1 - y_true) * np.square(y_pred) + y_true * np.square(np.maximum(margin - y_pred, 0)))


def whiten(batch, rms=0.038021):
    """Whiten a batch: per-sample mean removal, then one batch-global scale rms / sqrt(mean(batch**2))
    (voicemap/utils.py:88-101; the scale is taken over the whole un-centred batch)."""
    if len(batch.shape) != 3:
        raise ValueError('Input must be a 3D array of shape (n_segments, n_timesteps, 1).')
    sample_wise_mean = batch.mean(axis=1, keepdims=True)
    sample_wise_rescaling = rms / np.sqrt(np.power(batch, 2).mean())
    return (batch - sample_wise_mean) * sample_wise_rescaling


def _class_means(embeddings, n, k):
    return embeddings.reshape(k, n, -1).mean(axis=1)


def n_shot_task_evaluation(model, dataset, preprocessor, num_tasks, n, k, network_type='siamese',
                           distance='euclidean'):
    """Evaluate a network on k-way, n-shot classification tasks (voicemap/utils.py:104-216).

    Returns the number of correctly solved tasks.  A task is correct when the closest support item / class mean
    is index 0 (by construction of ``dataset.build_n_shot_task``)."""
    n_correct = 0

    if n == 1 and network_type == 'siamese':
        # Directly use siamese network to get pairwise verification score, minimum is closest
        for i_eval in tqdm(range(num_tasks)):
            query_sample, support_set_samples = dataset.build_n_shot_task(k, n)
            input_1 = np.stack([query_sample[0]] * k)[:, :, np.newaxis]
            input_2 = support_set_samples[0][:, :, np.newaxis]
            # Pass an empty list to the labels parameter as preprocessor functions work on batches not samples
            ([input_1, input_2], _) = preprocessor(([input_1, input_2], []))
            pred = model.predict([input_1, input_2])
            if np.argmin(pred[:, 0]) == 0:
                n_correct += 1
    elif n > 1 or network_type == 'classifier':
        # Create encoder network from earlier layers
        if network_type == 'siamese':
            encoder = model.layers[2]
        elif network_type == 'classifier':
            encoder = clone_model(model)
            encoder.set_weights(model.get_weights())
            encoder.pop()
        else:
            raise ValueError('mode must be one of (siamese, classifier)')

        for i_eval in tqdm(range(num_tasks)):
            query_sample, support_set_samples = dataset.build_n_shot_task(k, n)
            query_instance = preprocessor.instance_preprocessor(query_sample[0].reshape(1, -1, 1))
            support_set_instances = preprocessor.instance_preprocessor(support_set_samples[0][:, :, np.newaxis])

            query_embedding = encoder.predict(query_instance)
            support_set_embeddings = encoder.predict(support_set_instances)

            if distance == 'euclidean':
                # mean position of support set embeddings; labels are [class_1]*n + ... + [class_k]*n
                mean_support_set_embeddings = _class_means(support_set_embeddings, n, k)
                pred = np.sqrt(np.power(query_embedding - mean_support_set_embeddings, 2).sum(axis=1))
            elif distance == 'cosine':
                magnitudes = np.linalg.norm(support_set_embeddings, axis=1, keepdims=True)
                unit_vectors = support_set_embeddings / magnitudes
                mean_units = _class_means(unit_vectors, n, k)
                q = query_embedding[0]
                # scipy.spatial.distance.cdist(..., 'cosine')
                pred = 1.0 - (mean_units @ q) / (np.linalg.norm(mean_units, axis=1) * np.linalg.norm(q))
            elif distance == 'dot_product':
                magnitudes = np.linalg.norm(support_set_embeddings, axis=1, keepdims=True)
                unit_vectors = support_set_embeddings / magnitudes
                mean_units = _class_means(unit_vectors, n, k)
                mean_magnitudes = magnitudes.reshape(k, n).sum(axis=1, keepdims=True) / n
                pred = -np.dot(query_embedding[0, :][np.newaxis, :], (mean_magnitudes * mean_units).T)
            else:
                raise ValueError('Distance must be in (euclidean, cosine, dot_product)')

            if np.argmin(pred) == 0:
                n_correct += 1
    else:
        raise ValueError("n must be >= 1")

    return n_correct


def n_shot_task_evaluation_batched(model, dataset, preprocessor, num_tasks, n, k, network_type='siamese',
                                   distance='euclidean', tasks_per_launch=64):
    """Same tasks, same preprocessing and same decision rule as ``n_shot_task_evaluation`` but the encoder runs
    once per ``tasks_per_launch`` tasks instead of 1-2 tiny ``predict`` calls per task (the reference spends its
    per-epoch evaluation in 500 batch-5 forwards: experiments/train_siamese.py:75-78, voicemap/utils.py:121-137).
    Eval-mode embeddings do not depend on batch composition, so the count of correct tasks is identical for an
    identical task sequence (tasks are drawn in the same order, one ``build_n_shot_task`` call each)."""
    import torch
    from .engine import pair_head_loss
    siamese_direct = (n == 1 and network_type == 'siamese')
    if siamese_direct:
        encoder = model.encoder
    elif network_type == 'siamese':
        encoder = model.layers[2]
    elif network_type == 'classifier':
        encoder = clone_model(model)
        encoder.set_weights(model.get_weights())
        encoder.pop()
    else:
        raise ValueError('mode must be one of (siamese, classifier)')
    if n < 1:
        raise ValueError("n must be >= 1")
    if distance not in ('euclidean', 'cosine', 'dot_product'):
        raise ValueError('Distance must be in (euclidean, cosine, dot_product)')

    n_correct = 0
    done = 0
    while done < num_tasks:
        t = min(tasks_per_launch, num_tasks - done)
        queries, supports = [], []
        for _ in range(t):
            query_sample, support_set_samples = dataset.build_n_shot_task(k, n)
            if siamese_direct:
                # the reference whitens [query]*k and the k supports as two separate batches
                input_1 = np.stack([query_sample[0]] * k)[:, :, np.newaxis]
                input_2 = support_set_samples[0][:, :, np.newaxis]
                ([input_1, input_2], _) = preprocessor(([input_1, input_2], []))
                queries.append(input_1[:1])
                supports.append(input_2)
            else:
                queries.append(preprocessor.instance_preprocessor(query_sample[0].reshape(1, -1, 1)))
                supports.append(preprocessor.instance_preprocessor(support_set_samples[0][:, :, np.newaxis]))
        batch = np.concatenate(queries + supports, axis=0)
        xt = encoder._host_batch(batch)
        eng = encoder._get_engine()
        emb = eng.forward(xt.to(eng.device, non_blocking=True))
        eq, es = emb[:t], emb[t:].reshape(t, k * n, -1)
        if siamese_direct:
            w, b = model._head_device(eng.device)
            e1 = eq[:, None, :].expand(t, k, eq.shape[1]).reshape(t * k, -1).contiguous()
            prob, _, _ = pair_head_loss(e1, es.reshape(t * k, -1).contiguous(), w, b, model.distance_metric)
            n_correct += int((prob.reshape(t, k).argmin(dim=1) == 0).sum().item())
        else:
            eqn, esn = eq.cpu().numpy(), es.cpu().numpy()
            for i in range(t):
                support_set_embeddings = esn[i]
                if distance == 'euclidean':
                    means = _class_means(support_set_embeddings, n, k)
                    pred = np.sqrt(np.power(eqn[i:i + 1] - means, 2).sum(axis=1))
                else:
                    magnitudes = np.linalg.norm(support_set_embeddings, axis=1, keepdims=True)
                    mean_units = _class_means(support_set_embeddings / magnitudes, n, k)
                    if distance == 'cosine':
                        q = eqn[i]
                        pred = 1.0 - (mean_units @ q) / (np.linalg.norm(mean_units, axis=1) * np.linalg.norm(q))
                    else:
                        mean_magnitudes = magnitudes.reshape(k, n).sum(axis=1, keepdims=True) / n
                        pred = -np.dot(eqn[i][np.newaxis, :], (mean_magnitudes * mean_units).T)
                if np.argmin(pred) == 0:
                    n_correct += 1
        done += t
    return n_correct


class NShotEvaluationCallback(Callback):
    """Evaluate a network on n-shot classification tasks after every epoch (voicemap/utils.py:219-252)."""

    def __init__(self, num_tasks, n_shot, k_way, dataset, preprocessor=lambda x: x, mode='siamese', batch_tasks=0):
        super(NShotEvaluationCallback, self).__init__()
        self.batch_tasks = batch_tasks  # > 0: evaluate that many tasks per encoder launch (same result)
        self.num_tasks = num_tasks
        self.n_shot = n_shot
        self.k_way = k_way
        self.dataset = dataset
        self.preprocessor = preprocessor
        assert mode in ('siamese', 'classifier')
        self.mode = mode

    def on_epoch_end(self, epoch, logs=None):
        logs = logs if logs is not None else {}
        if self.batch_tasks > 0:
            n_correct = n_shot_task_evaluation_batched(self.model, self.dataset, self.preprocessor, self.num_tasks,
                                                       self.n_shot, self.k_way, network_type=self.mode,
                                                       tasks_per_launch=self.batch_tasks)
        else:
            n_correct = n_shot_task_evaluation(self.model, self.dataset, self.preprocessor, self.num_tasks,
                                               self.n_shot, self.k_way, network_type=self.mode)
        n_shot_acc = n_correct * 1. / self.num_tasks
        logs['val_{}-shot_acc'.format(self.n_shot)] = n_shot_acc
        print('val_{}-shot_acc: {:.4f}'.format(self.n_shot, n_shot_acc))
