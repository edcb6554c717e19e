"""Batch producers for ``fit_generator``: what Keras' ``workers=N, use_multiprocessing=True`` does for the reference
(experiments/train_siamese.py:71-72, train_classifier.py:127-128) -- several processes, each running its own copy of the
user's generator, feeding one queue -- rebuilt around shared memory so that a batch (2 x 64 x 48000 float64 = 49 MB
raw, 12 MB after the reference's preprocessing) is never pickled through a pipe.

* the parent draws one batch itself to learn the layout (it is also the first batch trained on), creates ``slots``
  shared-memory blocks of that size (+ 25 % head room) and forks ``workers`` children;
* a child reseeds ``numpy.random`` / ``random`` (forked copies would otherwise all produce the same "random" batches),
  then loops: take a free slot number, ``next(generator)``, copy every array of the batch into the slot, send
  ``(slot, structure)`` -- a few hundred bytes -- to the parent;
* the parent rebuilds the batch as numpy views of the slot (no copy).  A batch stays valid through the NEXT ``next()``
  call and is recycled by the one after it (the slot goes back to the children then); ``fit_generator`` has staged the
  arrays into its pinned buffers long before.  (Recycling at the very next call -- the first version -- made "look at
  the previous batch once more" a race with the producers, which one of this module's own tests lost now and then.)

A ``Sequence`` (experiments/train_classifier.py:48-86) is served the same way, except that the parent hands out the
indices of the epoch and returns the batches in index order (Keras' OrderedEnqueuer), and that the children are
forked anew every epoch because the reference's sequences reshuffle themselves in ``on_epoch_end``.

Children only run the user's generator (numpy, the FLAC decoder); they never touch CUDA, and they leave with
``os._exit`` so that no inherited finaliser does.  Order of batches across workers of a generator is arrival order, as
in Keras.  Python generators cannot be copied inside one process, so this needs ``fork``: on platforms without it, or
with ``workers <= 1`` / ``use_multiprocessing=False``, ``fit_generator`` keeps its single background thread.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import random
import traceback
import warnings
from multiprocessing import shared_memory

import numpy as np

_ALIGN = 64


def _flatten(batch, arrays):
    """Nested tuples / lists of arrays -> the same nesting with ('a', index) leaves; arrays collected in order."""
    if isinstance(batch, (list, tuple)):
        return ('l' if isinstance(batch, list) else 't', [_flatten(item, arrays) for item in batch])
    if isinstance(batch, np.ndarray):
        arrays.append(batch)
        return ('a', len(arrays) - 1)
    return ('o', batch)                                   # small picklable leaf (None, a scalar, ...)


def _rebuild(node, arrays):
    kind, value = node
    if kind == 'a':
        return arrays[value]
    if kind == 'o':
        return value
    items = [_rebuild(item, arrays) for item in value]
    return items if kind == 'l' else tuple(items)


def _layout(arrays):
    """[(offset, shape, dtype string)], total bytes: arrays packed back to back at 64-byte boundaries."""
    entries, offset = [], 0
    for a in arrays:
        entries.append((offset, a.shape, a.dtype.str))
        offset += -(-a.nbytes // _ALIGN) * _ALIGN
    return entries, offset


def _worker(index, count, seed, source, blocks, free_q, ready_q, task_q, stop):
    """Child process.  ``task_q`` is None for a generator (``next(source)`` per batch) or carries (sequence number,
    index) pairs for a Sequence (``source[index]``)."""
    code = 0
    try:
        np.random.seed((seed + 7919 * (index + 1)) % (2 ** 32))
        random.seed(seed + 104729 * (index + 1))
        from . import audio_io
        audio_io.DEFAULT_WORKERS = max(1, (os.cpu_count() or 1) // count)   # the decoder's threads, shared out
        iterator = iter(source) if task_q is None else None
        while not stop.is_set():
            slot = free_q.get()                            # a slot first, then work: see SequencePrefetcher
            if slot is None:
                break
            number = None
            if task_q is None:
                try:
                    batch = next(iterator)
                except StopIteration:
                    ready_q.put(('done', index, slot))
                    break
            else:
                task = task_q.get()
                if task is None:
                    free_q.put(slot)
                    break
                number, item = task
                batch = source[item]
            arrays = []
            tree = _flatten(batch, arrays)
            entries, total = _layout(arrays)
            block = blocks[slot]
            if total > block.size:
                ready_q.put(('error', index, slot, 'a batch of {} bytes does not fit the {}-byte slots sized from the '
                                                   'first batch'.format(total, block.size)))
                break
            for (offset, shape, dtype), a in zip(entries, arrays):
                np.ndarray(shape, dtype=dtype, buffer=block.buf, offset=offset)[...] = a
            ready_q.put(('batch', index, slot, tree, entries, number))
    except BaseException:                                  # surfaced to the training loop
        code = 1
        try:
            ready_q.put(('error', index, None, traceback.format_exc()))
        except Exception:
            pass
    finally:
        ready_q.close()
        ready_q.join_thread()                              # flush before the hard exit
        os._exit(code)


_LINGERING = []   # blocks whose mapping outlives their pool because a handed-out batch still views them


def _release(blocks):
    """Unlink the blocks' names now; unmap each block as soon as no numpy view of it is left.  A block that is still
    viewed (the last batch of an epoch, a batch kept by the caller, one shown in a traceback) stays mapped and is
    retried whenever another pool closes -- never unmapped under a live array."""
    for block in blocks:
        try:
            block.unlink()
        except FileNotFoundError:
            pass
    pending = _LINGERING + list(blocks)
    del _LINGERING[:]
    for block in pending:
        try:
            block.close()
        except BufferError:                                # exported pointers exist: a batch still views the block
            _LINGERING.append(block)


class _Pool:
    """Shared-memory slots, the queues and the forked children; the two prefetchers below differ in what they ask of
    them."""

    def __init__(self, source, workers, slots, first, seed, with_tasks):
        if 'fork' not in mp.get_all_start_methods():
            raise RuntimeError('multi-process batch production needs the fork start method')
        self.ctx = mp.get_context('fork')
        self.workers = int(workers)
        arrays = []
        _flatten(first, arrays)
        _, total = _layout(arrays)
        size = max(4096, int(total * 1.25))
        self.blocks = [shared_memory.SharedMemory(create=True, size=size) for _ in range(slots)]
        self.free_q = self.ctx.Queue()
        self.ready_q = self.ctx.Queue()
        self.task_q = self.ctx.Queue() if with_tasks else None
        self.stop = self.ctx.Event()
        for slot in range(slots):
            self.free_q.put(slot)
        if seed is None:
            seed = int(np.random.randint(0, 2 ** 31 - 1))  # reproducible under np.random.seed, distinct per worker
        self.procs = [self.ctx.Process(target=_worker, daemon=True,
                                       args=(i, self.workers, seed, source, self.blocks, self.free_q, self.ready_q, self.task_q,
                                             self.stop))
                      for i in range(self.workers)]
        with warnings.catch_warnings():
            # Python 3.12 warns that forking a multi-threaded process can deadlock the child.  The children here only run
            # the user's batch code and the io library (the arrangement of torch's DataLoader workers); they do not use
            # any lock a thread of the parent could hold, apart from the allocator's, which glibc re-initialises at fork.
            warnings.simplefilter("ignore", DeprecationWarning)
            for p in self.procs:
                p.start()
        self.closed = False

    def receive(self):
        """Next message of a child; raises if the children are gone or one of them failed."""
        while True:
            try:
                message = self.ready_q.get(timeout=5.0)
            except Exception:                              # queue.Empty: are the producers still there?
                if not any(p.is_alive() for p in self.procs):
                    raise RuntimeError('batch producer processes died without reporting an error') from None
                continue
            if message[0] == 'error':
                self.close()
                raise RuntimeError('batch producer {} failed:\n{}'.format(message[1], message[3]))
            return message

    def batch(self, message):
        _, _, slot, tree, entries, _ = message
        block = self.blocks[slot]
        # np.frombuffer (unlike np.ndarray(buffer=...)) keeps a buffer export of the mapping for as long as the array
        # lives, so close() cannot unmap a block under a batch somebody still holds (it defers, see _release)
        arrays = [np.frombuffer(block.buf, dtype=dtype, count=int(np.prod(shape, dtype=np.int64)), offset=offset)
                  .reshape(shape) for offset, shape, dtype in entries]
        return _rebuild(tree, arrays)

    def close(self):
        if self.closed:
            return
        self.closed = True
        self.stop.set()
        for _ in self.procs:
            self.free_q.put(None)
            if self.task_q is not None:
                self.task_q.put(None)
        for p in self.procs:
            p.join(timeout=0.5)
            if p.is_alive():
                p.terminate()
                p.join(timeout=2.0)
        for q in (self.free_q, self.ready_q, self.task_q):
            if q is not None:
                q.cancel_join_thread()
                q.close()
        _release(self.blocks)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ProcessPrefetcher:
    """``next()`` -> batches produced by ``workers`` forked copies of a generator, in arrival order."""

    def __init__(self, generator, workers, slots=None, seed=None):
        iterator = iter(generator)
        self._first = next(iterator)                       # also tells the slot size; trained on like any other batch
        # two slots stay with the caller (see the module docstring), the producers need at least one more
        self.pool = _Pool(iterator, workers, max(3, int(slots)) if slots else 2 * int(workers) + 2, self._first, seed,
                          False)
        self._held = []                                    # slots of the last two batches handed out (oldest first)
        self._live = self.pool.workers

    def next(self):
        if self._first is not None:
            batch, self._first = self._first, None
            return batch
        if len(self._held) == 2:                           # the batch before the previous one: recycle its slot
            self.pool.free_q.put(self._held.pop(0))
        while True:
            if self._live == 0:
                raise StopIteration
            message = self.pool.receive()
            if message[0] == 'done':
                self._live -= 1
                self.pool.free_q.put(message[2])
                continue
            self._held.append(message[2])
            return self.pool.batch(message)

    def close(self):
        self.pool.close()


class SequencePrefetcher:
    """Batches ``sequence[i]`` for a given list of indices, computed by ``workers`` forked children and handed out IN
    ORDER (Keras' OrderedEnqueuer).  One instance per epoch: the children hold the sequence as it was when they were
    forked, and the reference's sequences reshuffle themselves in ``on_epoch_end`` (experiments/train_classifier.py:83-86).

    No deadlock by construction: a child takes a slot before it takes a task, tasks leave the queue in order, and at
    most ``slots`` tasks are issued beyond the ones already consumed -- so the oldest unfinished task either owns a slot
    or finds one free."""

    def __init__(self, sequence, indices, workers, slots=None, seed=None):
        self.indices = list(indices)
        self.slots = max(3, int(slots)) if slots else 2 * int(workers) + 2
        self._first = sequence[self.indices[0]] if self.indices else None
        self.pool = _Pool(sequence, workers, self.slots, self._first, seed, True) if len(self.indices) > 1 else None
        self.issued = 1                                    # index 0 was computed here
        self.consumed = 0
        self.waiting = {}                                  # sequence number -> message, arrived early
        self._held = []                                    # slots of the last two batches handed out

    def _issue(self):
        # tasks beyond the consumed ones never exceed the slots the children can actually get: two stay with the caller
        while self.issued < len(self.indices) and self.issued < self.consumed + max(1, self.slots - 2):
            self.pool.task_q.put((self.issued, self.indices[self.issued]))
            self.issued += 1

    def next(self):
        if self.consumed >= len(self.indices):
            raise StopIteration
        if self.consumed == 0:
            self.consumed = 1
            if self.pool is not None:
                self._issue()
            return self._first
        if len(self._held) == 2:
            self.pool.free_q.put(self._held.pop(0))
        self._issue()
        while self.consumed not in self.waiting:
            message = self.pool.receive()
            self.waiting[message[5]] = message
        message = self.waiting.pop(self.consumed)
        self.consumed += 1
        self._held.append(message[2])
        return self.pool.batch(message)

    def close(self):
        if self.pool is not None:
            self.pool.close()
