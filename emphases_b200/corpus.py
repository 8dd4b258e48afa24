"""Native corpus reader / writer (csrc/corpus_io.cu) behind
from_files_to_files: all (TextGrid, wav) pairs are parsed on a C++ thread
pool straight into one pinned int16 audio buffer and one float64 word-time
array; files the native reader does not understand fall back to the Python
loaders."""
import ctypes
import os

import numpy as np
import torch

from . import _lib, scheduler


def _paths(paths):
    encoded = [os.fsencode(str(path)) for path in paths]
    array = (ctypes.c_char_p * len(encoded))(*encoded)
    return array, encoded          # keep `encoded` alive with the array


class Corpus:
    """Parsed corpus; use as a context manager"""

    def __init__(self, text_files, audio_files, threads=None):
        lib = _lib.load()
        self.lib = lib
        self.count = len(text_files)
        self.threads = threads or min(32, os.cpu_count() or 1)
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
        if self.handle:
            self.lib.emph_corpus_close(self.handle)
            self.handle = None

    def error(self, index):
        return self.lib.emph_corpus_error(self.handle, index).decode('utf-8', 'replace')

    def usable(self, sample_rate=16000):
        """Files the fast path can take: parsed, at the model's sample rate"""
        return (self.status == 0) & (self.sample_rate == sample_rate) & (self.n_words > 0)

    def load(self, mask, pin=True):
        """(indices, word-time arrays, PackedAudio[int16]) of the files
        selected by `mask` (compact: entry j belongs to file indices[j])"""
        mask = np.asarray(mask, dtype=bool) & (self.status == 0)
        indices = np.nonzero(mask)[0]
        lengths = self.n_samples[indices]
        offsets, total = scheduler.PackedAudio.layout(lengths)
        words = self.n_words[indices].astype(np.int64)
        word_offsets = np.concatenate([[0], np.cumsum(words[:-1])]) \
            if len(indices) else np.zeros(0, dtype=np.int64)
        buffer = torch.zeros(
            total, dtype=torch.int16,
            pin_memory=pin and torch.cuda.is_available())
        times = np.zeros((int(words.sum()), 2), dtype=np.float64)
        # the native fill visits every parsed file: files outside the mask
        # (e.g. another sample rate) get a scratch destination
        others = (self.status == 0) & ~mask
        scratch_audio = np.zeros(
            max(int(self.n_samples[others].max(initial=0)), 1), dtype=np.int16)
        scratch_times = np.zeros(
            (max(int(self.n_words[others].max(initial=0)), 1), 2))
        base_audio, base_times = buffer.data_ptr(), times.ctypes.data
        sample_offsets = np.full(
            self.count, (scratch_audio.ctypes.data - base_audio) // 2, dtype=np.int64)
        time_offsets = np.full(
            self.count, (scratch_times.ctypes.data - base_times) // 16, dtype=np.int64)
        sample_offsets[indices] = offsets
        time_offsets[indices] = word_offsets
        status = self.lib.emph_corpus_fill(
            self.handle, ctypes.c_void_p(base_audio),
            sample_offsets.ctypes.data, ctypes.c_void_p(base_times),
            time_offsets.ctypes.data, self.threads)
        if status != 0:
            raise _lib.EmphasesB200Error('emph_corpus_fill failed (short read)')
        per_file = [times[o:o + w] for o, w in zip(word_offsets, words)]
        return indices, per_file, scheduler.PackedAudio(buffer, offsets, lengths)

    def write_textgrids(self, output_paths, mask):
        encoded = [
            os.fsencode(str(path)) if m else b''
            for path, m in zip(output_paths, mask)]
        array = (ctypes.c_char_p * len(encoded))(*encoded)
        if self.lib.emph_corpus_write_textgrids(self.handle, array, self.threads):
            raise OSError('could not write some TextGrid files')


def write_scores(paths, scores, threads=None):
    """torch.save(scores[i], paths[i]) for a list of (1, W) float32 CPU tensors
    (emphases/core.py:112,177) on the native thread pool: the files are the zip
    archives torch.load reads, written without the Python pickler."""
    import numpy as np
    import torch
    lib = _lib.load()
    counts = np.array([int(score.shape[-1]) for score in scores], dtype=np.int32)
    offsets = np.concatenate([[0], np.cumsum(counts[:-1])]).astype(np.int64) \
        if len(counts) else np.zeros(0, dtype=np.int64)
    flat = torch.cat([
        score.detach().reshape(-1).to(device='cpu', dtype=torch.float32)
        for score in scores]) if len(scores) else torch.zeros(0)
    flat = np.ascontiguousarray(flat.numpy())
    encoded = [os.fsencode(str(path)) for path in paths]
    array = (ctypes.c_char_p * len(encoded))(*encoded)
    threads = threads or min(32, os.cpu_count() or 1)
    if lib.emph_write_score_files(
        array, flat.ctypes.data, offsets.ctypes.data, counts.ctypes.data,
        len(encoded), threads
    ):
        raise OSError('could not write some score files')
