"""LibriSpeech batcher behind the reference's ``voicemap.librispeech`` interface (constructor, ``__getitem__``,
pair / verification-batch / n-shot-task builders, ``.df`` columns: voicemap/librispeech.py:31-281).

Host-side I/O, outside the CUDA hot path.  The implementation is array based: after the index is loaded the
corpus is three numpy vectors (speaker id, length in samples, sex flag) plus a path list; every draw is a
vectorised ``np.random`` call on those vectors instead of DataFrame merges, and the DataFrame the reference's callers
read (``dataset.df``) is kept only as a view of the same rows.

Sampling distributions are the reference's:
  * files are drawn without replacement with probability proportional to their length (`df.sample(weights='length')`,
    voicemap/librispeech.py:145,157,161,219,226,236);
  * an alike pair is a uniform pick among all (anchor, same-speaker file) combinations of 2n length-weighted anchors,
    the anchor itself included (the merge-then-sample of voicemap/librispeech.py:144-151);
  * a differing pair is a length-weighted file plus a length-weighted file of any speaker absent from the first draw
    (voicemap/librispeech.py:157-165);
  * labels of a verification batch are 0 (same speaker) for the first half and 1 for the second (:179-194).

Decoding: the reference calls ``soundfile.read`` (voicemap/librispeech.py:104,267); the default here is
``audio_io.read`` -- our own C FLAC decoder (``csrc/vm_flac.c``, bit-exact, no dependency), ``wave`` for .wav,
``soundfile`` for anything else if it is installed.  With the default reader a corpus is indexed from the FLAC stream
headers (4 KB per file) instead of decoding every utterance, and the clips of a batch are decoded on a thread pool
(the decoder runs outside the GIL); random fragment offsets are still drawn in batch order on the calling thread.

Extensions (keyword-only): ``reader`` plugs in any decoder ``path -> (samples, rate)``, ``index`` injects a ready index
table so the batcher runs without the corpus on disk, ``data_path`` replaces ``config.PATH``, ``decode_workers`` sizes
the thread pool (default: up to 16, 0/1 = decode on the calling thread).
"""
from __future__ import annotations

import os
from collections import defaultdict

import numpy as np
import pandas as pd

from . import audio_io
from .config import LIBRISPEECH_SAMPLING_RATE, PATH
from .keras_compat import Sequence

sex_to_label = {'M': False, 'F': True}
label_to_sex = {flag: sex for sex, flag in sex_to_label.items()}

_LABEL_KINDS = ('sex', 'speaker')
_SPEAKER_FIELDS = ('id', 'sex', 'subset', 'minutes', 'name')
_FILE_FIELDS = ('id', 'filepath', 'length', 'seconds')


def _file_length(path, reader):
    """Samples in an utterance: from the FLAC header when our decoder is the reader, else by decoding the file."""
    if reader is audio_io.read and path.lower().endswith('.flac'):
        frames = audio_io.flac_info(path)['frames']
        if frames:
            return frames
    samples, _ = reader(path)
    return len(samples)


def read_speaker_table(path):
    """Parse LibriSpeech's SPEAKERS.TXT (``;`` comment lines, then ``ID | SEX | SUBSET | MINUTES | NAME`` records).
    Records that do not have exactly five fields are dropped, which is what the reference's
    ``read_csv(..., error_bad_lines=False)`` does with them (voicemap/librispeech.py:64-71)."""
    records = []
    with open(path, encoding='utf-8', errors='replace') as handle:
        for line in handle:
            if line.startswith(';') or not line.strip():
                continue
            parts = [p.strip() for p in line.rstrip('\n').split('|')]
            if len(parts) != len(_SPEAKER_FIELDS):
                continue
            records.append((int(parts[0]), parts[1], parts[2], float(parts[3]), parts[4]))
    return pd.DataFrame.from_records(records, columns=_SPEAKER_FIELDS)


def _flac_files(subset_dir):
    """(speaker id, path) of every .flac below ``subset_dir`` (layout <subset>/<speaker>/<chapter>/<utterance>.flac)."""
    found = []
    for folder, _, names in os.walk(subset_dir):
        flacs = sorted(n for n in names if n.endswith('.flac'))
        if flacs:
            speaker = int(os.path.basename(os.path.dirname(os.path.normpath(folder))))
            found.extend((speaker, os.path.join(folder, n)) for n in flacs)
    return found


class LibriSpeechDataset(Sequence):
    """``dataset[i] -> (float64[fragment_length], label)`` plus verification batches and k-way n-shot tasks.

    # Arguments (same meaning as voicemap/librispeech.py:31-41)
        subsets: subset name or list of names ('dev-clean', 'train-clean-100', ...)
        seconds: fragment length; files not longer than this are dropped unless ``pad``
        label: 'speaker' or 'sex'
        stochastic: random fragment position (and random split of the padding) instead of the file start
        pad: keep short files and zero-pad them to ``seconds``
        cache: reuse / write ``<data_path>/data/<subset>.index.csv``
    """

    def __init__(self, subsets, seconds, label='speaker', stochastic=True, pad=False, cache=True, *,
                 reader=None, index=None, data_path=None, decode_workers=None):
        assert label in _LABEL_KINDS, "Label type must be one of ('sex', 'speaker')"
        self.subset = subsets
        self.label = label
        self.stochastic = stochastic
        self.pad = pad
        self.fragment_seconds = seconds
        self.fragment_length = int(seconds * LIBRISPEECH_SAMPLING_RATE)
        self.reader = reader if reader is not None else audio_io.read
        self.decode_workers = decode_workers
        self.data_path = data_path if data_path is not None else PATH

        print('Initialising LibriSpeechDataset with minimum length = {}s and subsets = {}'.format(seconds, subsets))
        names = [subsets] if isinstance(subsets, str) else list(subsets)
        if index is None:
            table = pd.concat([self._subset_table(name, cache) for name in names], ignore_index=True)
        else:
            table = pd.DataFrame(index)
        if not pad:
            table = table[table['seconds'] > seconds]
        self.unique_speakers = int(table['id'].nunique())

        # row number == dataset id; the speaker's LibriVox id moves to `speaker_id` (voicemap/librispeech.py:88-96)
        table = table.rename(columns={'id': 'speaker_id', 'minutes': 'speaker_minutes'}).reset_index(drop=True)
        table['id'] = np.arange(len(table))

        # per-file facts, indexed by dataset id for the lifetime of the object
        self._paths = table['filepath'].tolist()
        self._speaker = table['speaker_id'].to_numpy()
        self._weight = table['length'].to_numpy(dtype=np.float64)
        self._sex = table['sex'].tolist()
        self.df = table          # property: also derives the sampling structures from the rows of the frame
        # the reference's lookup tables, kept for code that pokes at them
        self.datasetid_to_filepath = dict(enumerate(self._paths))
        self.datasetid_to_speaker_id = dict(enumerate(self._speaker.tolist()))
        self.datasetid_to_sex = dict(enumerate(self._sex))
        print('Finished indexing data. {} usable files found.'.format(len(self)))

    # ------------------------------------------------------------------------------------------------ index
    def _subset_table(self, name, cache):
        csv_path = os.path.join(self.data_path, 'data', '{}.index.csv'.format(name))
        if cache and os.path.exists(csv_path):
            return pd.read_csv(csv_path)
        speakers = read_speaker_table(os.path.join(self.data_path, 'data', 'LibriSpeech', 'SPEAKERS.TXT'))
        files = pd.DataFrame(self.index_subset(name, reader=self.reader, data_path=self.data_path),
                             columns=_FILE_FIELDS)
        table = speakers.merge(files, on='id')
        table.to_csv(csv_path, index=False)
        return table

    @staticmethod
    def index_subset(subset, reader=None, data_path=None):
        """One record {id, filepath, length, seconds} per .flac of ``subset`` (voicemap/librispeech.py:243-281);
        lengths come from the FLAC headers with the default reader, from a full decode with a custom one."""
        from tqdm import tqdm
        reader = reader if reader is not None else audio_io.read
        root = os.path.join(data_path if data_path is not None else PATH, 'data', 'LibriSpeech', subset)
        print('Indexing {}...'.format(subset))
        records = []
        for speaker, path in tqdm(_flac_files(root)):
            length = _file_length(path, reader)
            records.append({'id': speaker, 'filepath': path, 'length': length,
                            'seconds': length / float(LIBRISPEECH_SAMPLING_RATE)})
        return records

    # ------------------------------------------------------------------------------------------------ items
    @property
    def df(self):
        """The index table the reference's callers read -- and replace: experiments/wide_vs_tall.py:55-78 deep-copies
        a dataset and assigns a frame with fewer rows (same ``id`` values) to train on fewer speakers.  Assigning
        restricts every draw to the rows of the new frame, in its row order, as ``self.df.sample`` does there."""
        return self._df

    @df.setter
    def df(self, frame):
        self._df = frame
        self._active = frame['id'].to_numpy(dtype=np.int64)
        members = defaultdict(list)
        for row in self._active:
            members[self._speaker[row]].append(row)
        self._members = {speaker: np.asarray(rows) for speaker, rows in members.items()}
        self._group_size = np.zeros(len(self._paths), dtype=np.int64)
        self._group_size[self._active] = [len(self._members[self._speaker[row]]) for row in self._active]

    def __len__(self):
        return len(self._active)

    def num_classes(self):
        return len(self._members)

    def _fragment(self, samples):
        """Crop to ``fragment_length`` (random offset if stochastic) and, with ``pad``, zero-fill short clips --
        randomly split front/back if stochastic, at the back otherwise (voicemap/librispeech.py:105-124)."""
        want = self.fragment_length
        start = np.random.randint(0, max(len(samples) - want, 1)) if self.stochastic else 0
        piece = samples[start:start + want]
        missing = want - len(piece)
        if missing <= 0 or not self.pad:
            return piece
        lead = np.random.randint(0, missing) if self.stochastic else 0
        out = np.zeros(want, dtype=piece.dtype)
        out[lead:lead + len(piece)] = piece
        return out

    def _label(self, index):
        if self.label == 'speaker':
            label = self._speaker[index]
        elif self.label == 'sex':
            label = sex_to_label[self._sex[index]]
        else:
            raise ValueError("Label type must be one of ('sex', 'speaker')")
        return label

    def _native_flac(self, index):
        return self.reader is audio_io.read and self._paths[index].lower().endswith('.flac')

    def _draw_fragment(self, length):
        """The random draws of ``_fragment`` for a file of ``length`` samples, in the same order, without the audio:
        (start, samples to take, leading zeros)."""
        want = self.fragment_length
        start = np.random.randint(0, max(length - want, 1)) if self.stochastic else 0
        take = max(min(want, length - start), 0)
        missing = want - take
        if missing <= 0 or not self.pad:
            return start, take, 0
        return start, take, (np.random.randint(0, missing) if self.stochastic else 0)

    def _place(self, piece, take, lead, path):
        if len(piece) != take:
            raise ValueError('{}: the index promises {} more samples than the file holds; delete the cached '
                             '*.index.csv and re-index'.format(path, take - len(piece)))
        if take == self.fragment_length or not self.pad:
            return piece
        out = np.zeros(self.fragment_length, dtype=piece.dtype)
        out[lead:lead + take] = piece
        return out

    def __getitem__(self, index):
        if self._native_flac(index):      # decode only the frames under the fragment
            start, take, lead = self._draw_fragment(int(self._weight[index]))
            piece, _ = audio_io.read_flac_range(self._paths[index], start, take)
            return self._place(piece, take, lead, self._paths[index]), self._label(index)
        samples, _ = self.reader(self._paths[index])
        return self._fragment(samples), self._label(index)

    def _native_clips(self, rows):
        """(len(rows), fragment_length) float64 batch straight from the decoder library, or None when this dataset
        cannot use it (custom reader, non-FLAC files, or clips that come out shorter than the fragment length).  The
        crop offsets and padding splits are drawn here, in row order, exactly as ``__getitem__`` draws them."""
        if not all(self._native_flac(r) for r in rows):
            return None
        state = np.random.get_state()
        plans = [self._draw_fragment(int(self._weight[r])) for r in rows]
        if not self.pad and any(take != self.fragment_length for _, take, _ in plans):
            np.random.set_state(state)           # ragged clips: the per-item path handles them (and redraws)
            return None
        starts, takes, leads = zip(*plans) if plans else ((), (), ())
        try:
            return audio_io.read_fragments([self._paths[r] for r in rows], starts, takes, self.fragment_length, leads,
                                           workers=self.decode_workers)
        except audio_io.AudioDecodeError as exc:
            if 'fewer samples' in str(exc):
                raise ValueError('{}; delete the cached *.index.csv and re-index'.format(exc)) from None
            raise

    def _items(self, rows):
        """``[self[r] for r in rows]`` with the files decoded concurrently.  Every random draw happens on the calling
        thread in row order -- before decoding when the file lengths are known from the index (FLAC + our decoder: only
        the fragment is decoded, by the library's own threads), after it otherwise -- so the random stream is the one
        the serial loop consumes."""
        rows = [int(r) for r in rows]
        batch = self._native_clips(rows)
        if batch is not None:
            return [(batch[i], self._label(r)) for i, r in enumerate(rows)]
        workers = self.decode_workers
        if all(self._native_flac(r) for r in rows):
            plans = [self._draw_fragment(int(self._weight[r])) for r in rows]
            jobs = [(self._paths[r], start, take) for r, (start, take, _) in zip(rows, plans)]
            pieces = audio_io.read_many(jobs, reader=audio_io.read_flac_range, workers=workers)
            return [(self._place(piece, take, lead, self._paths[r]), self._label(r))
                    for r, (_, take, lead), (piece, _) in zip(rows, plans, pieces)]
        if workers is None and self.reader is not audio_io.read:
            workers = 1                      # a user-supplied Python reader gains nothing from threads
        decoded = audio_io.read_many([self._paths[r] for r in rows], reader=self.reader, workers=workers)
        return [(self._fragment(samples), self._label(r)) for r, (samples, _) in zip(rows, decoded)]

    def _clips(self, rows):
        rows = [int(r) for r in rows]
        batch = self._native_clips(rows)
        if batch is not None:
            return batch
        return np.stack([clip for clip, _ in self._items(rows)])

    # ------------------------------------------------------------------------------------------------ draws
    def _draw(self, count, among=None):
        """``count`` distinct dataset ids, probability proportional to file length, optionally restricted to the id
        array ``among``.  (numpy's weighted choice without replacement is what the pandas 0.23 pinned by the reference
        calls; pandas >= 2.2 refuses the same request when count * max(weight) > sum(weights), which small corpora
        hit at the reference's batch sizes.)"""
        pool = self._active if among is None else np.asarray(among)
        w = self._weight[pool]
        return pool[np.random.choice(len(pool), size=count, replace=False, p=w / w.sum())]

    def get_alike_pairs(self, num_pairs):
        """``num_pairs`` (id, id) tuples whose files share a speaker (voicemap/librispeech.py:143-153)."""
        anchors = self._draw(2 * num_pairs)
        ends = np.cumsum(self._group_size[anchors])          # combinations contributed by each anchor
        picks = np.random.choice(int(ends[-1]), size=num_pairs, replace=False)
        which = np.searchsorted(ends, picks, side='right')
        nth = picks - (ends[which] - self._group_size[anchors[which]])
        first = anchors[which]
        second = [self._members[self._speaker[a]][j] for a, j in zip(first, nth)]
        return list(zip(first.tolist(), (int(s) for s in second)))

    def get_differing_pairs(self, num_pairs):
        """``num_pairs`` (id, id) tuples from different speakers (voicemap/librispeech.py:155-167)."""
        first = self._draw(num_pairs)
        others = self._active[~np.isin(self._speaker[self._active], self._speaker[first])]
        second = self._draw(num_pairs, among=others)
        return list(zip(first.tolist(), second.tolist()))

    def build_verification_batch(self, batchsize):
        """([input_1, input_2], labels): inputs (B, T, 1), labels (B, 1); first half alike pairs labelled 0, second
        half differing pairs labelled 1 (voicemap/librispeech.py:169-196)."""
        half = batchsize // 2
        pairs = self.get_alike_pairs(half) + self.get_differing_pairs(half)
        left, right = zip(*pairs)
        inputs = [self._clips(side)[:, :, np.newaxis] for side in (left, right)]
        labels = np.repeat([0.0, 1.0], half)[:, np.newaxis]
        return inputs, labels

    def yield_verification_batches(self, batchsize):
        """Endless stream of verification batches (voicemap/librispeech.py:198-202)."""
        while True:
            yield self.build_verification_batch(batchsize)

    def build_n_shot_task(self, k, n=1):
        """((query, label), (support[k*n, T], labels[k*n])): the first n support clips are other files of the query's
        speaker, followed by n clips of each of k-1 other speakers (voicemap/librispeech.py:204-240)."""
        if k >= self.unique_speakers:
            raise ValueError('k must be smaller than the number of unique speakers in this dataset!')
        if k <= 1:
            raise ValueError('k must be greater than or equal to one!')
        query = int(self._draw(1)[0])
        query_sample = self[query]
        speaker = self._speaker[query]
        same = self._members[speaker]
        support = [self._draw(n, among=same[same != query])]
        rivals = np.random.choice(np.asarray([s for s in self._members if s != speaker]), k - 1, replace=False)
        support.extend(self._draw(n, among=self._members[r]) for r in rivals)
        items = self._items(np.concatenate(support))
        clips, labels = zip(*items)
        return query_sample, (np.stack(clips), np.stack(labels))
