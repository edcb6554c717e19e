"""voicemap.librispeech batcher with the reference's signatures (voicemap/librispeech.py:15-281), ported to
Python 3.  Host-side I/O only (FLAC decoding is outside the hot path): the audio reader is pluggable --
``soundfile`` when installed, else any callable ``reader(path) -> (samples, samplerate)`` -- and an index
DataFrame can be injected so the sampling logic runs without LibriSpeech on disk.
"""
from __future__ import annotations

import os

import numpy as np
import pandas as pd
from tqdm import tqdm

from .config import LIBRISPEECH_SAMPLING_RATE, PATH
from .keras_compat import Sequence

sex_to_label = {'M': False, 'F': True}
label_to_sex = {False: 'M', True: 'F'}


def _default_reader(path):
    try:
        import soundfile as sf
    except ImportError as exc:  # pragma: no cover - depends on the environment
        raise ImportError("reading LibriSpeech FLAC files needs `soundfile`; pass reader=... to "
                          "LibriSpeechDataset to plug in another decoder") from exc
    return sf.read(path)


def _sample_by_length(df, n):
    """``df.sample(n, weights='length')`` with the semantics of the pandas the reference pins (0.23): n distinct rows
    drawn with numpy's weighted choice without replacement from the global ``np.random`` state.  pandas >= 2.2 refuses
    that draw whenever n * max(weight) > sum(weights) ("Weighted sampling cannot be achieved with replace=False"),
    which small corpora hit at the reference's batch sizes."""
    w = df['length'].to_numpy(dtype=np.float64)
    locs = np.random.choice(len(df), size=n, replace=False, p=w / w.sum())
    return df.iloc[locs]


class LibriSpeechDataset(Sequence):
    """Sequence whose __getitem__ returns (raw audio fragment float64[fragment_length], label); also builds
    verification batches and k-way n-shot tasks.

    # Arguments (voicemap/librispeech.py:31)
        subsets, seconds, label ('speaker'|'sex'), stochastic, pad, cache: as in the reference.
    # Extensions (keyword-only)
        reader: callable(path) -> (samples, samplerate).
        index: pre-built pandas DataFrame with the reference's index columns
               (id, sex, subset, minutes, name, filepath, length, seconds); skips disk indexing.
        data_path: root containing ``data/LibriSpeech`` (default config.PATH).
    """
    def __init__(self, subsets, seconds, label='speaker', stochastic=True, pad=False, cache=True, *,
                 reader=None, index=None, data_path=None):
        assert label in ('sex', 'speaker'), 'Label type must be one of (\'sex\', \'speaker\')'
        self.subset = subsets
        self.fragment_seconds = seconds
        self.fragment_length = int(seconds * LIBRISPEECH_SAMPLING_RATE)
        self.stochastic = stochastic
        self.pad = pad
        self.label = label
        self.reader = reader or _default_reader
        self.data_path = data_path or PATH

        print('Initialising LibriSpeechDataset with minimum length = {}s and subsets = {}'.format(seconds, subsets))

        if isinstance(subsets, str):
            subsets = [subsets]

        if index is not None:
            self.df = index.copy()
        else:
            cached_df = []
            found_cache = {s: False for s in subsets}
            if cache:
                for s in subsets:
                    subset_index_path = self.data_path + '/data/{}.index.csv'.format(s)
                    if os.path.exists(subset_index_path):
                        cached_df.append(pd.read_csv(subset_index_path))
                        found_cache[s] = True

            if all(found_cache.values()) and cache:
                self.df = pd.concat(cached_df)
            else:
                df = pd.read_csv(self.data_path + '/data/LibriSpeech/SPEAKERS.TXT', skiprows=11, delimiter='|',
                                 on_bad_lines='skip')
                df.columns = [col.strip().replace(';', '').lower() for col in df.columns]
                df = df.assign(
                    sex=df['sex'].apply(lambda x: x.strip()),
                    subset=df['subset'].apply(lambda x: x.strip()),
                    name=df['name'].apply(lambda x: x.strip()),
                )
                audio_files = []
                for subset, found in found_cache.items():
                    if not found:
                        audio_files += self.index_subset(subset, reader=self.reader, data_path=self.data_path)
                df = pd.merge(df, pd.DataFrame(audio_files))
                self.df = pd.concat(cached_df + [df])

            for s in subsets:
                self.df[self.df['subset'] == s].to_csv(self.data_path + '/data/{}.index.csv'.format(s), index=False)

        # Trim too-small files
        if not self.pad:
            self.df = self.df[self.df['seconds'] > self.fragment_seconds]
        self.unique_speakers = len(self.df['id'].unique())

        # Renaming for clarity
        self.df = self.df.rename(columns={'id': 'speaker_id', 'minutes': 'speaker_minutes'})

        # Index of dataframe has direct correspondence to item in dataset
        self.df = self.df.reset_index(drop=True)
        self.df = self.df.assign(id=self.df.index.values)

        self.datasetid_to_filepath = self.df.to_dict()['filepath']
        self.datasetid_to_speaker_id = self.df.to_dict()['speaker_id']
        self.datasetid_to_sex = self.df.to_dict()['sex']

        print('Finished indexing data. {} usable files found.'.format(len(self)))

    def __getitem__(self, index):
        instance, samplerate = self.reader(self.datasetid_to_filepath[index])
        # Choose a random sample of the file
        if self.stochastic:
            fragment_start_index = np.random.randint(0, max(len(instance) - self.fragment_length, 1))
        else:
            fragment_start_index = 0

        instance = instance[fragment_start_index:fragment_start_index + self.fragment_length]

        # Check for required length and pad if necessary
        if self.pad and len(instance) < self.fragment_length:
            less_timesteps = self.fragment_length - len(instance)
            if self.stochastic:
                # random number of 0s before, the rest after
                before_len = np.random.randint(0, less_timesteps)
                after_len = less_timesteps - before_len
                instance = np.pad(instance, (before_len, after_len), 'constant')
            else:
                instance = np.pad(instance, (0, less_timesteps), 'constant')

        if self.label == 'sex':
            label = sex_to_label[self.datasetid_to_sex[index]]
        elif self.label == 'speaker':
            label = self.datasetid_to_speaker_id[index]
        else:
            raise ValueError('Label type must be one of (\'sex\', \'speaker\')')

        return instance, label

    def __len__(self):
        return len(self.df)

    def num_classes(self):
        return len(self.df['speaker_id'].unique())

    def get_alike_pairs(self, num_pairs):
        """List of 2-tuples of dataset IDs belonging to the same speaker (voicemap/librispeech.py:143-153)."""
        alike_pairs = pd.merge(
            _sample_by_length(self.df, num_pairs * 2),
            self.df,
            on='speaker_id'
        ).sample(num_pairs)[['speaker_id', 'id_x', 'id_y']]
        return list(zip(alike_pairs['id_x'].values, alike_pairs['id_y'].values))

    def get_differing_pairs(self, num_pairs):
        """List of 2-tuples of dataset IDs belonging to different speakers (voicemap/librispeech.py:155-167)."""
        random_sample = _sample_by_length(self.df, num_pairs)
        random_sample_from_other_speakers = _sample_by_length(
            self.df[~self.df['speaker_id'].isin(random_sample['speaker_id'])], num_pairs)
        return list(zip(random_sample['id'].values, random_sample_from_other_speakers['id'].values))

    def build_verification_batch(self, batchsize):
        """Batch of verification pairs: first half same-speaker pairs (label 0), second half different speakers
        (label 1) (voicemap/librispeech.py:169-196).  Returns ([input_1, input_2] each (B, T, 1), labels (B, 1))."""
        half = batchsize // 2
        alike_pairs = self.get_alike_pairs(half)
        input_1_alike = np.stack([self[i][0] for i in list(zip(*alike_pairs))[0]])
        input_2_alike = np.stack([self[i][0] for i in list(zip(*alike_pairs))[1]])

        differing_pairs = self.get_differing_pairs(half)
        input_1_different = np.stack([self[i][0] for i in list(zip(*differing_pairs))[0]])
        input_2_different = np.stack([self[i][0] for i in list(zip(*differing_pairs))[1]])

        input_1 = np.vstack([input_1_alike, input_1_different])[:, :, np.newaxis]
        input_2 = np.vstack([input_2_alike, input_2_different])[:, :, np.newaxis]

        outputs = np.append(np.zeros(half), np.ones(half))[:, np.newaxis]

        return [input_1, input_2], outputs

    def yield_verification_batches(self, batchsize):
        """Convenience function to yield verification batches forever."""
        while True:
            ([input_1, input_2], labels) = self.build_verification_batch(batchsize)
            yield ([input_1, input_2], labels)

    def build_n_shot_task(self, k, n=1):
        """k-way n-shot task: (query_sample, support_set_samples); the first n support samples belong to the
        query's speaker (voicemap/librispeech.py:204-240)."""
        if k >= self.unique_speakers:
            raise ValueError('k must be smaller than the number of unique speakers in this dataset!')

        if k <= 1:
            raise ValueError('k must be greater than or equal to one!')

        query = _sample_by_length(self.df, 1)
        query_sample = self[query.index.values[0]]

        is_query_speaker = self.df['speaker_id'] == query['speaker_id'].values[0]
        not_same_sample = self.df.index != query.index.values[0]
        correct_samples = _sample_by_length(self.df[is_query_speaker & not_same_sample], n)

        # Sample k-1 speakers
        other_support_set_speakers = np.random.choice(
            self.df[~is_query_speaker]['speaker_id'].unique(), k - 1, replace=False)

        other_support_samples = []
        for i in range(k - 1):
            is_same_speaker = self.df['speaker_id'] == other_support_set_speakers[i]
            other_support_samples.append(
                _sample_by_length(self.df[~is_query_speaker & is_same_speaker], n)
            )
        support_set = pd.concat([correct_samples] + other_support_samples)
        support_set_samples = tuple(np.stack(i) for i in zip(*[self[i] for i in support_set.index]))

        return query_sample, support_set_samples

    @staticmethod
    def index_subset(subset, reader=None, data_path=None):
        """Index a subset: speaker ID, filepath and length of every .flac (voicemap/librispeech.py:243-281)."""
        reader = reader or _default_reader
        data_path = data_path or PATH
        audio_files = []
        print('Indexing {}...'.format(subset))
        subset_len = 0
        for root, folders, files in os.walk(data_path + '/data/LibriSpeech/{}/'.format(subset)):
            subset_len += len([f for f in files if f.endswith('.flac')])

        progress_bar = tqdm(total=subset_len)
        for root, folders, files in os.walk(data_path + '/data/LibriSpeech/{}/'.format(subset)):
            if len(files) == 0:
                continue
            librispeech_id = int(root.split('/')[-2])
            for f in files:
                if not f.endswith('.flac'):
                    continue
                progress_bar.update(1)
                instance, samplerate = reader(os.path.join(root, f))
                audio_files.append({
                    'id': librispeech_id,
                    'filepath': os.path.join(root, f),
                    'length': len(instance),
                    'seconds': len(instance) * 1. / LIBRISPEECH_SAMPLING_RATE
                })
        progress_bar.close()
        return audio_files
