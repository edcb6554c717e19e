"""Runs the REFERENCE'S batcher (/root/reference/voicemap/librispeech.py, imported unchanged from where it lies) and
this repo's batcher side by side on the same LibriSpeech tree and prints a JSON verdict.  Used by
tests/test_reference_suite.py in the build container; a separate process so that the import shims stay out of the test
process.  Shims (environment only): `soundfile` -> voicemap_b200.audio_io, `keras.utils.Sequence` -> a bare class.
The reference's indexing branch is Python-2 only (`dict.iteritems`, pandas-0.23 `error_bad_lines`), so the index CSV is
written by OUR indexer first and the reference then loads it through its cache branch
(voicemap/librispeech.py:48-59) -- which also shows that the two index formats are interchangeable.

usage: python run_reference_batcher.py <data root holding data/LibriSpeech/...>
"""
import importlib.util
import json
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    data_root = sys.argv[1]
    os.environ["VOICEMAP_PATH"] = data_root
    from voicemap_b200 import audio_io
    from voicemap_b200.librispeech import LibriSpeechDataset as Ours

    keras = types.ModuleType("keras")
    keras.utils = types.ModuleType("keras.utils")
    keras.utils.Sequence = type("Sequence", (), {})
    sys.modules.update({"keras": keras, "keras.utils": keras.utils, "soundfile": audio_io})
    spec = importlib.util.spec_from_file_location("reference_librispeech", "/root/reference/voicemap/librispeech.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    assert ref.PATH == data_root, (ref.PATH, data_root)

    verdict = {"items": 0, "mismatches": []}
    Ours("dev-clean", 1.0, cache=True)                              # writes <root>/data/dev-clean.index.csv
    for seconds, stochastic, pad, label in [(3, False, False, "speaker"), (3, True, False, "speaker"),
                                            (3.05, True, True, "speaker"), (3.05, False, True, "sex"),
                                            (1, True, False, "sex")]:
        theirs = ref.LibriSpeechDataset("dev-clean", seconds, label=label, stochastic=stochastic, pad=pad)
        ours = Ours("dev-clean", seconds, label=label, stochastic=stochastic, pad=pad)
        key = "{}s stochastic={} pad={} label={}".format(seconds, stochastic, pad, label)
        if len(theirs) != len(ours) or theirs.num_classes() != ours.num_classes():
            verdict["mismatches"].append(key + ": sizes")
            continue
        columns = ["speaker_id", "filepath", "length", "seconds", "id", "sex", "subset"]
        if not theirs.df[columns].reset_index(drop=True).equals(ours.df[columns].reset_index(drop=True)):
            verdict["mismatches"].append(key + ": index tables differ")
        for i in range(len(ours)):
            np.random.seed(1000 + i)
            a, la = theirs[i]
            state_theirs = np.random.get_state()[1][:8].tolist(), np.random.get_state()[2]
            np.random.seed(1000 + i)
            b, lb = ours[i]
            state_ours = np.random.get_state()[1][:8].tolist(), np.random.get_state()[2]
            same = a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a, b) and la == lb
            if not same or state_theirs != state_ours:                 # same clip AND same random numbers consumed
                verdict["mismatches"].append("{}: item {}".format(key, i))
            verdict["items"] += 1
        # a batch drawn through the threaded fragment path consumes the stream like the reference's serial loop
        rows = list(range(0, len(ours), 7))
        np.random.seed(5)
        serial = [theirs[i][0] for i in rows]
        np.random.seed(5)
        batch = ours._clips(rows)
        if not all(np.array_equal(x, y) for x, y in zip(serial, batch)):
            verdict["mismatches"].append(key + ": batch")
    # pair and task draws under the same seed.  pandas' DataFrame.sample draws from numpy's global RandomState, as
    # our array-based draws do, so the sequences can be compared element by element, not just in distribution.
    # A draw that the reference cannot make (too few files of a speaker, ...) must fail on our side too; draws that
    # only pandas 3 refuses ("Weighted sampling cannot be achieved", the reference pins pandas 0.23) are skipped.
    def both(label, theirs_call, ours_call, seed, same):
        results = []
        for call in (theirs_call, ours_call):
            np.random.seed(seed)
            try:
                results.append(("ok", call()))
            except ValueError as exc:
                results.append(("refused" if "Weighted sampling cannot" in str(exc) else "error", None))
        if results[0][0] == "refused":
            verdict["skipped"] = verdict.get("skipped", 0) + 1
        elif results[0][0] != results[1][0] or (results[0][0] == "ok" and not same(results[0][1], results[1][1])):
            verdict["mismatches"].append("{}, seed {} ({} / {})".format(label, seed, results[0][0], results[1][0]))
        elif results[0][0] == "error":
            verdict["both_failed"] = verdict.get("both_failed", 0) + 1

    def same_pairs(a, b):
        return [(int(i), int(j)) for i, j in a] == [(int(i), int(j)) for i, j in b]

    def same_task(a, b):
        (q1, l1), (s1, sl1) = a
        (q2, l2), (s2, sl2) = b
        return np.array_equal(q1, q2) and l1 == l2 and np.array_equal(s1, s2) and np.array_equal(sl1, sl2)

    verdict["draws"] = 0
    theirs = ref.LibriSpeechDataset("dev-clean", 3, stochastic=True)
    ours = Ours("dev-clean", 3, stochastic=True)
    for seed in range(40):
        both("alike pairs", lambda: list(theirs.get_alike_pairs(16)), lambda: ours.get_alike_pairs(16), seed, same_pairs)
        both("differing pairs", lambda: list(theirs.get_differing_pairs(16)), lambda: ours.get_differing_pairs(16), seed,
             same_pairs)
        for k, n in ((5, 1), (4, 2), (3, 3)):
            both("{}-way {}-shot task".format(k, n), lambda: theirs.build_n_shot_task(k, n),
                 lambda: ours.build_n_shot_task(k, n), seed, same_task)
        verdict["draws"] += 1

    # a reduced index assigned from outside, as experiments/wide_vs_tall.py:55-78 does (fewer speakers, fewer files, the
    # original ids, a shuffled row order): both batchers must keep drawing the same pairs and tasks from the subset
    verdict["reduced"] = 0
    keep = theirs.df[theirs.df["speaker_id"].isin(sorted(theirs.df["speaker_id"].unique())[::2])]
    keep = keep[keep["id"] % 5 != 0].sample(frac=1.0, random_state=3)
    theirs.df = keep
    ours.df = ours.df.loc[keep.index]
    if len(theirs) != len(ours) or theirs.num_classes() != ours.num_classes():
        verdict["mismatches"].append("reduced index: sizes")
    inside = set(int(i) for i in keep["id"])

    def same_pairs_inside(a, b):
        return same_pairs(a, b) and set(int(i) for pair in b for i in pair) <= inside

    for seed in range(20):
        both("reduced index: alike pairs", lambda: list(theirs.get_alike_pairs(8)), lambda: ours.get_alike_pairs(8), seed,
             same_pairs_inside)
        both("reduced index: differing pairs", lambda: list(theirs.get_differing_pairs(8)),
             lambda: ours.get_differing_pairs(8), seed, same_pairs_inside)
        both("reduced index: task", lambda: theirs.build_n_shot_task(4, 2), lambda: ours.build_n_shot_task(4, 2), seed,
             same_task)
        verdict["reduced"] += 1
    print(json.dumps(verdict))


if __name__ == "__main__":
    main()
