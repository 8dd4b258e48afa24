"""Word alignment container + Praat TextGrid I/O.

Stand-in for the slice of `pypar.Alignment` the reference's hot path uses
(emphases/core.py:49,107,111,366-400): len(), [int] -> word with
start()/end()/duration(), [slice] -> alignment re-based to t = 0,
word_bounds(sample_rate, hopsize, silences=True), save(file).  pypar itself is
a third-party dependency that is not vendored in the reference; any object
with that duck-typed surface (including a real pypar.Alignment) is accepted
by the API functions through `as_times`.
"""
import os
import re

import numpy as np

SILENCE = '<silent>'


class Word:

    def __init__(self, word, start, end):
        self.word = word
        self._start = float(start)
        self._end = float(end)

    def __str__(self):
        return self.word

    def __repr__(self):
        return f'Word({self.word!r}, {self._start}, {self._end})'

    def start(self):
        return self._start

    def end(self):
        return self._end

    def duration(self):
        return self._end - self._start


class Alignment:

    def __init__(self, source):
        """source: a .TextGrid path, a list of Word, or a list of
        (label, start, end) tuples"""
        if isinstance(source, (str, bytes, os.PathLike)):
            self._words = _fill_gaps(read_textgrid(source))
        else:
            self._words = [
                word if isinstance(word, Word) else Word(*word)
                for word in source]

    @classmethod
    def from_times(cls, times, labels=None):
        return cls([
            Word(f'w{i}' if labels is None else labels[i], start, end)
            for i, (start, end) in enumerate(times)])

    def __len__(self):
        return len(self._words)

    def __iter__(self):
        return iter(self._words)

    def __getitem__(self, index):
        if isinstance(index, slice):
            words = self._words[index]
            origin = words[0].start() if words else 0.
            return Alignment([
                Word(word.word, word.start() - origin, word.end() - origin)
                for word in words])
        return self._words[index]

    def __str__(self):
        return ' '.join(
            str(word) for word in self._words if str(word) != SILENCE)

    def start(self):
        return self._words[0].start()

    def end(self):
        return self._words[-1].end()

    def duration(self):
        return self.end() - self.start()

    def times(self):
        """(W, 2) float64 array of (start, end) seconds (cached)"""
        cached = getattr(self, '_times', None)
        if cached is None or len(cached) != len(self._words):
            cached = np.array(
                [(word.start(), word.end()) for word in self._words],
                dtype=np.float64).reshape(-1, 2)
            self._times = cached
        return cached

    def word_bounds(self, sample_rate, hopsize=1, silences=False):
        return [
            (int(word.start() * sample_rate / hopsize),
             int(word.end() * sample_rate / hopsize))
            for word in self._words
            if silences or str(word) != SILENCE]

    def save(self, file):
        write_textgrid(file, self._words)


def as_times(alignment):
    """(W, 2) float64 word times from any duck-typed alignment"""
    if hasattr(alignment, 'times'):
        return alignment.times()
    if isinstance(alignment, np.ndarray):
        return alignment.astype(np.float64).reshape(-1, 2)
    return np.array(
        [(alignment[i].start(), alignment[i].end())
         for i in range(len(alignment))],
        dtype=np.float64).reshape(-1, 2)


###############################################################################
# TextGrid I/O (long and short Praat text formats)
###############################################################################


def _fill_gaps(words, tolerance=1e-9):
    """Insert silence words into gaps so the words tile [0, end]"""
    result = []
    cursor = 0.
    for word in words:
        label = word.word
        if label.strip() in ('', 'sp', 'sil'):
            label = SILENCE
        if word.start() - cursor > tolerance:
            result.append(Word(SILENCE, cursor, word.start()))
        result.append(Word(label, word.start(), word.end()))
        cursor = word.end()
    return result


def read_textgrid(file):
    """Words of the 'words' interval tier (or the first interval tier)"""
    with open(file, 'rb') as stream:
        raw = stream.read()
    if raw[:2] in (b'\xff\xfe', b'\xfe\xff'):
        text = raw.decode('utf-16')
    else:
        text = raw.decode('utf-8-sig')
    tiers = _parse_tiers(text)
    if not tiers:
        raise ValueError(f'No interval tier found in {file}')
    for name, intervals in tiers:
        if name.lower() in ('words', 'word'):
            return intervals
    return tiers[0][1]


_NUMBER = r'[-+]?\d+(?:\.\d*)?(?:[eE][-+]?\d+)?'


def _parse_tiers(text):
    # Tokenise into quoted strings and numbers; works for long and short form.
    # Quoted strings are matched first so digits inside labels survive; the
    # bracketed indices of the long form ("intervals [3]:") are dropped.
    text = re.sub(r'"(?:[^"]|"")*"|\[\s*\d*\s*\]',
                  lambda m: m.group(0) if m.group(0).startswith('"') else ' ',
                  text)
    tokens = re.findall(r'"((?:[^"]|"")*)"|(' + _NUMBER + r')', text)
    items = [
        ('s', s.replace('""', '"')) if n == '' else ('n', float(n))
        for s, n in tokens]
    tiers = []
    i = 0
    while i < len(items):
        kind, value = items[i]
        if kind == 's' and value in ('IntervalTier', 'TextTier'):
            name = items[i + 1][1]
            count = int(items[i + 4][1])
            cursor = i + 5
            intervals = []
            if value == 'IntervalTier':
                for _ in range(count):
                    xmin, xmax, label = (
                        items[cursor][1], items[cursor + 1][1],
                        items[cursor + 2][1])
                    intervals.append(Word(label, xmin, xmax))
                    cursor += 3
                tiers.append((name, intervals))
            else:
                cursor += 2 * count
            i = cursor
        else:
            i += 1
    return tiers


def write_textgrid(file, words):
    xmax = words[-1].end() if words else 0.
    lines = [
        'File type = "ooTextFile"',
        'Object class = "TextGrid"',
        '',
        'xmin = 0',
        f'xmax = {xmax!r}',
        'tiers? <exists>',
        'size = 2',
        'item []:']
    for index, name in enumerate(('words', 'phones'), 1):
        lines += [
            f'    item [{index}]:',
            '        class = "IntervalTier"',
            f'        name = "{name}"',
            '        xmin = 0',
            f'        xmax = {xmax!r}',
            f'        intervals: size = {len(words)}']
        for j, word in enumerate(words, 1):
            label = str(word).replace('"', '""')
            if label == SILENCE:
                label = 'sp' if name == 'words' else 'sil'
            lines += [
                f'        intervals [{j}]:',
                f'            xmin = {word.start()!r}',
                f'            xmax = {word.end()!r}',
                f'            text = "{label}"']
    with open(file, 'w', encoding='utf-8') as stream:
        stream.write('\n'.join(lines) + '\n')
