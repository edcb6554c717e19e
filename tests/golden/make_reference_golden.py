"""Generates tests/golden/reference_executed.npz by EXECUTING THE REFERENCE'S OWN SOURCE (read from /root/reference at
run time, never copied): voicemap/models.py is imported unchanged and voicemap/utils.py with its single Python-2 print
statement rewritten in memory, against the numpy stand-in for Keras in keras_standin.py (this image has no Keras /
TensorFlow).  See keras_standin.py for what this pins (architecture, wiring, heads, loss, preprocessing, n-shot rules:
the reference's code) and what it does not (the arithmetic inside each Keras layer: restated).

Run in the build container, from the repo root:  python tests/golden/make_reference_golden.py
The GPU box has no /root/reference; tests read only the committed fixture.
"""
import importlib.util
import os
import re
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import keras_standin  # noqa: E402
import reference_cases as RC  # noqa: E402

REFERENCE = "/root/reference"


def load_reference():
    keras_standin.install()
    spec = importlib.util.spec_from_file_location("ref_voicemap_models", os.path.join(REFERENCE, "voicemap/models.py"))
    models = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(models)
    path = os.path.join(REFERENCE, "voicemap/utils.py")
    source = open(path).read()
    source, n = re.subn(r"^(\s*)print (.+)$", r"\1print(\2)", source, flags=re.M)   # py2 print statement(s)
    assert n == 1, n
    utils = types.ModuleType("ref_voicemap_utils")
    exec(compile(source, path, "exec"), utils.__dict__)
    return models, utils


def block_outputs(encoder, x):
    """Activations after each MaxPool of the reference-built Sequential (first positions only, to keep the file small)."""
    outs = []
    for layer in encoder.layers:
        x = layer.compute(x)
        if isinstance(layer, keras_standin.MaxPool1D):
            outs.append(x[:, :24, :].copy())
    return outs


def perturbed(encoder, rel, seed):
    """Same encoder with every weight moved by `rel` relative noise (is a decision count robust to GPU-size errors?)."""
    rng = np.random.default_rng(seed)
    clone = keras_standin.clone_model(encoder)
    clone.set_weights([w * (1 + rel * rng.standard_normal(w.shape)) for w in encoder.get_weights()])
    return clone


def main():
    models, utils = load_reference()
    out = {}
    encoders = {}
    for name in RC.ENCODER_CASES:
        filters, emb, n, length, _, _ = RC.ENCODER_CASES[name]
        params, x = RC.encoder_inputs(name)
        encoder = models.get_baseline_convolutional_encoder(filters, emb)              # voicemap/models.py:6
        encoder.set_weights(RC.keras_weight_list(params))
        encoders[name] = encoder
        kinds = [type(layer).__name__ for layer in encoder.layers]
        out[f"{name}_layer_kinds"] = np.array(kinds)
        out[f"{name}_emb"] = encoder.predict(x)
        for i, act in enumerate(block_outputs(encoder, x.astype(np.float64)), start=1):
            out[f"{name}_block{i}_head"] = act

        half = n // 2
        scale = float(np.median(np.linalg.norm(out[f"{name}_emb"][:half] - out[f"{name}_emb"][half:], axis=1)))
        out[f"{name}_head_scale"] = scale
        y = (np.arange(half) % 2).astype(np.float64)[:, None]
        for metric in ("uniform_euclidean", "weighted_l1"):
            siamese = models.build_siamese_net(encoder, (length, 1), distance_metric=metric)   # voicemap/models.py:44
            assert siamese.layers[2] is encoder and len(siamese.layers) == 6                  # voicemap/utils.py:141
            kernel, bias = RC.head_weights(metric, emb, scale)
            siamese.layers[-1].set_weights([kernel, bias])
            prob = siamese.predict([x[:half], x[half:]])
            out[f"{name}_{metric}_prob"] = prob
            out[f"{name}_{metric}_contrastive"] = utils.contrastive_loss(y, prob)              # voicemap/utils.py:77
        for metric in ("weighted_euclidean", "uniform_l1", "dot_product", "cosine_distance"):
            try:
                models.build_siamese_net(encoder, (length, 1), distance_metric=metric)
                raise SystemExit("expected NotImplementedError")
            except NotImplementedError:
                pass
        try:
            models.build_siamese_net(encoder, (length, 1), distance_metric="manhattan")
            raise SystemExit("expected AssertionError")
        except AssertionError:
            pass

    # classifier as experiments/train_classifier.py:110-112 builds it
    filters, emb, n, length, _, _ = RC.ENCODER_CASES["f16"]
    params, x = RC.encoder_inputs("f16")
    classifier = models.get_baseline_convolutional_encoder(filters, emb, (length, 1))
    classifier.add(keras_standin.Dense(7, activation="softmax"))
    kernel, bias = RC.classifier_head(emb)
    classifier.set_weights(RC.keras_weight_list(params) + [kernel, bias])
    out["f16_classifier_prob"] = classifier.predict(x)

    # preprocessing, voicemap/utils.py:22-34 and :88-101
    raw = RC.raw_clips()
    out["preprocess_ds4"] = utils.preprocess_instances(4)(raw)
    out["preprocess_ds4_no_whiten"] = utils.preprocess_instances(4, whitening=False)(raw)
    out["preprocess_ds1"] = utils.preprocess_instances(1)(raw)
    out["whiten"] = utils.whiten(raw)
    pre = utils.BatchPreProcessor("siamese", utils.preprocess_instances(4))
    (a, b), labels = pre(([raw, raw[::-1]], np.array([[0.0], [1.0], [1.0]])))
    out["batch_preprocessor_left"], out["batch_preprocessor_right"], out["batch_preprocessor_labels"] = a, b, labels

    # k-way n-shot evaluation, voicemap/utils.py:104-216, on deterministic task streams
    encoder = encoders["f16"]
    emb_dim = RC.ENCODER_CASES["f16"][1]
    # head scale for the pairwise case: typical query-support embedding distance, so that the sigmoid does not saturate
    # (a saturated head returns 1.0 for every pair and argmin then "solves" every task by returning index 0)
    (query, _), (support, _) = RC.TaskDataset(99).build_n_shot_task(5, 1)
    instance = utils.preprocess_instances(RC.NSHOT_DOWNSAMPLING)
    spread = np.linalg.norm(encoder.predict(instance(query.reshape(1, -1, 1))) -
                            encoder.predict(instance(support[:, :, None])), axis=1)
    nshot_scale = float(np.median(spread))
    out["nshot_head_scale"] = nshot_scale
    counts, robust = [], []
    for network_type, n_shot, k_way, distance, tasks, seed, noise in RC.NSHOT_CASES:
        def build(enc):
            if network_type == "siamese":
                model = models.build_siamese_net(enc, (RC.TaskDataset(0).samples // RC.NSHOT_DOWNSAMPLING, 1))
                model.layers[-1].set_weights(list(RC.head_weights("uniform_euclidean", emb_dim, nshot_scale)))
                return model
            model = keras_standin.Sequential(list(enc.layers))
            model.add(keras_standin.Dense(7, activation="softmax"))
            model.layers[-1].set_weights(list(RC.classifier_head(emb_dim)))
            return model

        def run(enc):
            preprocessor = utils.BatchPreProcessor("siamese", utils.preprocess_instances(RC.NSHOT_DOWNSAMPLING))
            return utils.n_shot_task_evaluation(build(enc), RC.TaskDataset(seed, noise=noise), preprocessor, tasks, n_shot, k_way,
                                                network_type=network_type, distance=distance)
        count = run(encoder)
        counts.append(count)
        robust.append(all(run(perturbed(encoder, 3e-5, s)) == count for s in range(4)))
        print(network_type, n_shot, k_way, distance, "->", count, "of", tasks, "robust" if robust[-1] else "FRAGILE")
    out["nshot_correct"] = np.array(counts)
    out["nshot_robust"] = np.array(robust)

    np.savez_compressed(RC.FIXTURE, **out)
    print("wrote", RC.FIXTURE, os.path.getsize(RC.FIXTURE), "bytes")


if __name__ == "__main__":
    main()
