"""Native corpus reader / writer (csrc/corpus_io.cu) behind
from_files_to_files: all (TextGrid, wav) pairs are parsed on a C++ thread
pool straight into one pinned int16 audio buffer and one float64 word-time
array; files the native reader does not understand fall back to the Python
loaders."""
import ctypes
import os
import threading

import numpy as np
import torch

from . import _lib, scheduler


def _encode(path):
    # (os.fsencode costs ~1 us per path through its fspath / type checks; a
    # corpus has tens of thousands)
    if type(path) is str:
        return path.encode('utf-8', 'surrogateescape')
    return os.fsencode(path)


def _blob(paths, suffix=''):
    """All paths (str) + suffix as one buffer of NUL-terminated strings"""
    if not paths:
        return b''
    separator = suffix + '\0'
    return (separator.join(paths) + separator).encode('utf-8', 'surrogateescape')


def _paths(paths):
    encoded = [_encode(path) for path in paths]
    array = (ctypes.c_char_p * len(encoded))(*encoded)
    return array, encoded          # keep `encoded` alive with the array


def _staging(samples, pinned):
    return scheduler.staging(samples, torch.int16, pinned)


class Corpus:
    """Parsed corpus; use as a context manager"""

    def __init__(self, text_files, audio_files, threads=None):
        lib = _lib.load()
        self.lib = lib
        self.count = len(text_files)
        self.threads = threads or min(32, os.cpu_count() or 1)
        if all(type(path) is str for path in text_files) and \
                all(type(path) is str for path in audio_files):
            self.handle = lib.emph_corpus_open_blob(
                _blob(text_files), _blob(audio_files), self.count, self.threads)
        else:
            text_array, self._text_keep = _paths(text_files)
            audio_array, self._audio_keep = _paths(audio_files)
            self.handle = lib.emph_corpus_open(
                text_array, audio_array, self.count, self.threads)
        n = max(self.count, 1)
        self.status = np.zeros(n, dtype=np.int32)
        self.sample_rate = np.zeros(n, dtype=np.int32)
        self.channels = np.zeros(n, dtype=np.int32)
        self.n_samples = np.zeros(n, dtype=np.int64)
        self.n_words = np.zeros(n, dtype=np.int32)
        lib.emph_corpus_info(
            self.handle, self.status.ctypes.data, self.sample_rate.ctypes.data,
            self.channels.ctypes.data, self.n_samples.ctypes.data,
            self.n_words.ctypes.data)
        for array in ('status', 'sample_rate', 'channels', 'n_samples', 'n_words'):
            setattr(self, array, getattr(self, array)[:self.count])

    def __enter__(self):
        return self

    def __exit__(self, *args):
        self.close()

    def close(self):
        worker = getattr(self, '_fill_worker', None)
        if worker is not None:
            worker.join()                 # the native reader still uses the handle
            self._fill_worker = None
        if self.handle:
            self.lib.emph_corpus_close(self.handle)
            self.handle = None

    def error(self, index):
        return self.lib.emph_corpus_error(self.handle, index).decode('utf-8', 'replace')

    def usable(self, sample_rate=16000):
        """Files the fast path can take: parsed (16-bit PCM + TextGrid), at
        `sample_rate` (None: any rate)"""
        mask = (self.status == 0) & (self.n_words > 0)
        if sample_rate is not None:
            mask &= self.sample_rate == sample_rate
        return mask

    def load(self, mask, pin=True, group_samples=1 << 26):
        """(indices, word-time arrays, PackedAudio[int16]) of the files
        selected by `mask` (compact: entry j belongs to file indices[j]).

        Decoding runs on a background thread, group by group (about
        `group_samples` samples each, in file order): the returned
        PackedAudio's `ready(j)` blocks until entries 0..j are resident, so
        the first launches upload while later files are still being read."""
        mask = np.asarray(mask, dtype=bool) & (self.status == 0)
        indices = np.nonzero(mask)[0]
        lengths = self.n_samples[indices]
        offsets, total = scheduler.PackedAudio.layout(lengths)
        words = self.n_words[indices].astype(np.int64)
        word_offsets = np.concatenate([[0], np.cumsum(words[:-1])]) \
            if len(indices) else np.zeros(0, dtype=np.int64)
        buffer = _staging(total, pin and torch.cuda.is_available())
        times = np.zeros((int(words.sum()), 2), dtype=np.float64)
        sample_offsets = np.zeros(self.count, dtype=np.int64)
        time_offsets = np.zeros(self.count, dtype=np.int64)
        sample_offsets[indices] = offsets
        time_offsets[indices] = word_offsets
        per_file = [times[o:o + w] for o, w in zip(word_offsets, words)]
        packed = scheduler.PackedAudio(buffer, offsets, lengths)
        if not len(indices):
            return indices, per_file, packed

        # groups of consecutive entries, ~group_samples each
        ends = np.cumsum(lengths)
        cuts = np.searchsorted(
            ends, np.arange(group_samples, int(ends[-1]) + group_samples, group_samples))
        bounds = sorted(set([0] + [min(int(c) + 1, len(indices)) for c in cuts] + [len(indices)]))
        groups = [(a, b) for a, b in zip(bounds[:-1], bounds[1:]) if b > a]
        file_ids = indices.astype(np.int32)
        state = {'filled': 0, 'error': None}
        condition = threading.Condition()
        keep = (buffer, times, sample_offsets, time_offsets, file_ids)

        def fill():
            for first, last in groups:
                part = np.ascontiguousarray(file_ids[first:last])
                status = self.lib.emph_corpus_fill_files(
                    self.handle, part.ctypes.data, len(part),
                    ctypes.c_void_p(keep[0].data_ptr()), sample_offsets.ctypes.data,
                    ctypes.c_void_p(times.ctypes.data), time_offsets.ctypes.data,
                    self.threads)
                with condition:
                    if status != 0:
                        state['error'] = _lib.EmphasesB200Error(
                            'emph_corpus_fill_files failed (short read)')
                    state['filled'] = last
                    condition.notify_all()
                if status != 0:
                    return

        def ready(entry):
            with condition:
                condition.wait_for(
                    lambda: state['filled'] > entry or state['error'] is not None)
                if state['error'] is not None:
                    raise state['error']

        worker = threading.Thread(target=fill, daemon=True)
        worker.start()
        self._fill_worker = worker
        packed.ready = ready
        return indices, per_file, packed

    def write_textgrids_to_prefixes(self, prefixes, mask):
        """`{prefix}.TextGrid` for the files selected by `mask` (str prefixes)"""
        mask = np.ascontiguousarray(mask, dtype=np.uint8)
        chosen = [prefix for prefix, m in zip(prefixes, mask) if m]
        if self.lib.emph_corpus_write_textgrids_blob(
            self.handle, _blob(chosen, '.TextGrid'), mask.ctypes.data, self.threads
        ):
            raise OSError('could not write some TextGrid files')

    def write_textgrids(self, output_paths, mask):
        encoded = [
            _encode(path) if m else b''
            for path, m in zip(output_paths, mask)]
        array = (ctypes.c_char_p * len(encoded))(*encoded)
        if self.lib.emph_corpus_write_textgrids(self.handle, array, self.threads):
            raise OSError('could not write some TextGrid files')


def write_score_rows(paths, flat_scores, counts, threads=None, suffix=''):
    """torch.save(row_i[None], paths[i] + suffix) where row i is the next
    counts[i] values of ONE flat fp32 host tensor: no per-file tensor objects.
    threads < 0: that many threads of the call's own instead of the shared
    worker pool (a writer running beside a decode that occupies the pool)"""
    lib = _lib.load()
    flat_scores = flat_scores.detach()
    if flat_scores.device.type != 'cpu' or flat_scores.dtype != torch.float32 \
            or not flat_scores.is_contiguous():
        flat_scores = flat_scores.to(device='cpu', dtype=torch.float32).contiguous()
    counts = np.ascontiguousarray(counts, dtype=np.int32)
    if int(counts.sum()) != flat_scores.numel() or len(counts) != len(paths):
        raise ValueError('write_score_rows: counts do not match the scores')
    starts = np.concatenate([[0], np.cumsum(counts[:-1], dtype=np.int64)]) \
        if len(counts) else np.zeros(0, dtype=np.int64)
    threads = threads or min(32, os.cpu_count() or 1)
    starts = np.ascontiguousarray(starts, dtype=np.int64)
    if all(type(path) is str for path in paths):
        status = lib.emph_write_score_rows_blob(
            _blob(paths, suffix), flat_scores.data_ptr(), starts.ctypes.data,
            counts.ctypes.data, len(paths), threads)
    else:
        pointers = (flat_scores.data_ptr() + 4 * starts).astype(np.uint64)
        array, keep = _paths([f'{path}{suffix}' for path in paths])
        status = lib.emph_write_score_rows(
            array, pointers.ctypes.data, counts.ctypes.data, len(paths), threads)
        del keep
    if status:
        raise OSError('could not write some score files')


def write_scores(paths, scores, threads=None):
    """torch.save(scores[i], paths[i]) for a list of (1, W) float32 CPU tensors
    (emphases/core.py:112,177) on the native thread pool: the files are the zip
    archives torch.load reads, written without the Python pickler."""
    import numpy as np
    import torch
    lib = _lib.load()
    rows = [
        score.detach().reshape(-1) if (
            score.device.type == 'cpu' and score.dtype == torch.float32 and
            score.is_contiguous())
        else score.detach().reshape(-1).to(device='cpu', dtype=torch.float32).contiguous()
        for score in scores]
    counts = np.array([row.numel() for row in rows], dtype=np.int32)
    pointers = np.array([row.data_ptr() for row in rows], dtype=np.uint64)
    encoded = [os.fsencode(path) for path in paths]
    array = (ctypes.c_char_p * len(encoded))(*encoded)
    threads = threads or min(32, os.cpu_count() or 1)
    if lib.emph_write_score_rows(
        array, pointers.ctypes.data, counts.ctypes.data, len(encoded), threads
    ):
        raise OSError('could not write some score files')
