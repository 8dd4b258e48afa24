"""Import harness for the UNMODIFIED reference package (test infrastructure only).

ORACLE / TEST INFRASTRUCTURE -- never imported by the product path
(`emphases_b200/`).  Only `oracle/gen_golden.py` (run by hand in the build
container, where `/root/reference` exists) uses this file.

`import emphases` from /root/reference fails in this image because ten
third-party modules are absent.  None of them do
arithmetic on the hot path except `librosa.filters.mel` (restated in
`oracle/emphases_oracle.py:mel_basis`, following the published librosa
algorithm) and `pypar.Alignment` (restated here as `Alignment`, see the
"parity unpinned" note in `oracle/emphases_oracle.py`).  This module injects
minimal stand-ins into `sys.modules` so that the reference's own
`emphases.Model`, `emphases.preprocess`, `emphases.downsample`,
`emphases.segment`, `emphases.from_alignment_and_audio` and `emphases.loss`
run on CPU exactly as written.
"""
import contextlib
import copy
import importlib
import math
import os
import sys
import types

import numpy as np
import torch

REFERENCE_ROOT = os.environ.get('EMPHASES_REF', '/root/reference')

SILENCE = '<silent>'


###############################################################################
# pypar stand-in (third-party, un-vendored, unpinned: setup.py lists 'pypar')
###############################################################################


class Phoneme:

    def __init__(self, phoneme, start, end):
        self.phoneme = phoneme
        self._start = start
        self._end = end

    def __str__(self):
        return self.phoneme

    def start(self):
        return self._start

    def end(self):
        return self._end

    def duration(self):
        return self._end - self._start


class Word:
    """A word with start/end times in seconds (one phoneme spanning it)"""

    def __init__(self, word, phonemes):
        self.word = word
        self.phonemes = phonemes

    def __str__(self):
        return self.word

    def __len__(self):
        return len(self.phonemes)

    def start(self):
        return self.phonemes[0].start()

    def end(self):
        return self.phonemes[-1].end()

    def duration(self):
        return self.end() - self.start()


class Alignment:
    """Duck-typed stand-in for pypar.Alignment (emphases/core.py:49,366-400)

    Behaviour relied on by the reference hot path:
      len(), [int] -> Word, [slice] -> Alignment re-based so the first word
      starts at t=0, word_bounds(sr, hop, silences=True) ->
      [(int(start*sr/hop), int(end*sr/hop))].
    """

    def __init__(self, words):
        self._words = list(words)

    @classmethod
    def from_times(cls, times, labels=None):
        """times: iterable of (start, end) seconds"""
        words = []
        for i, (start, end) in enumerate(times):
            label = f'w{i}' if labels is None else labels[i]
            words.append(
                Word(label, [Phoneme(label, float(start), float(end))]))
        return cls(words)

    def __len__(self):
        return len(self._words)

    def __getitem__(self, idx):
        if isinstance(idx, slice):
            words = copy.deepcopy(self._words[idx])
            if words:
                origin = words[0].start()
                for word in words:
                    for phoneme in word.phonemes:
                        phoneme._start = phoneme._start - origin
                        phoneme._end = phoneme._end - origin
            return Alignment(words)
        return self._words[idx]

    def __iter__(self):
        return iter(self._words)

    def start(self):
        return self._words[0].start()

    def end(self):
        return self._words[-1].end()

    def duration(self):
        return self.end() - self.start()

    def word_bounds(self, sample_rate, hopsize=1, silences=False):
        words = [
            word for word in self._words
            if silences or str(word) != SILENCE]
        return [
            (int(word.start() * sample_rate / hopsize),
             int(word.end() * sample_rate / hopsize))
            for word in words]

    def save(self, file):
        raise NotImplementedError('stub')


###############################################################################
# Stub installation
###############################################################################


def _module(name, **attrs):
    module = types.ModuleType(name)
    for key, value in attrs.items():
        setattr(module, key, value)
    sys.modules[name] = module
    return module


def _notify(*args, **kwargs):
    def decorator(fn):
        return fn
    return decorator


def _iterator(iterable, message=None, initial=0, total=None):
    return iterable


def _checkpoint_load(file, model, optimizer=None, map_location='cpu'):
    """torchutil.checkpoint.load(file, model) -> (model, optimizer, state)"""
    state = torch.load(file, map_location=map_location, weights_only=False)
    model.load_state_dict(state['model'])
    if optimizer is not None and 'optimizer' in state:
        optimizer.load_state_dict(state['optimizer'])
    rest = {k: v for k, v in state.items() if k not in ('model', 'optimizer')}
    return model, optimizer, rest


class _Metric:
    """torchutil.metrics.Metric: the interface the reference subclasses"""

    def __init__(self, *args, **kwargs):
        self.reset()

    def update(self, *args, **kwargs):
        pass

    def reset(self):
        pass

    def __call__(self):
        return {}


# torchutil (third party, un-pinned, absent here) -- restated from its
# published metrics module: PARITY UNPINNED for these three classes; the
# reference's own subclasses (emphases/evaluate/metrics.py) run unmodified on
# top of them.
class _Average(_Metric):
    """Running average: update(values, count) adds values.sum() and count"""

    def __call__(self):
        return (self.total / self.count).item()

    def update(self, values, count):
        self.total += values.sum()
        self.count += count

    def reset(self):
        self.total = 0.
        self.count = 0


class _MeanStd(_Metric):
    """Welford running mean / sample standard deviation over python floats"""

    def __call__(self):
        return self.mean, math.sqrt(self.m2 / (self.count - 1))

    def update(self, values):
        for value in values:
            self.count += 1
            delta = value - self.mean
            self.mean += delta / self.count
            self.m2 += delta * (value - self.mean)

    def reset(self):
        self.m2 = 0.
        self.mean = 0.
        self.count = 0


class _PearsonCorrelation(_Metric):
    """Pearson correlation from precomputed means / standard deviations"""

    def __init__(self, predicted_mean, predicted_std, target_mean, target_std):
        self.reset()
        self.mean = predicted_mean
        self.std = predicted_std
        self.target_mean = target_mean
        self.target_std = target_std

    def __call__(self):
        return (
            1. / self.count *
            (self.total / (self.std * self.target_std))).item()

    def update(self, predicted, target):
        self.total += (
            (predicted - self.mean) * (target - self.target_mean)).sum()
        self.count += predicted.numel()

    def reset(self):
        self.count = 0
        self.total = 0.


def install(overrides=None):
    """Install stubs; `overrides` maps emphases config names -> values,
    applied the way yapecs applies a --config file (to config.defaults
    before `from .config.defaults import *`)."""
    from oracle import emphases_oracle

    overrides = dict(overrides or {})

    def configure(name, defaults_module):
        for key, value in overrides.items():
            setattr(defaults_module, key, value)

    _module('yapecs', configure=configure)
    _module('GPUtil', getGPUs=lambda: [])

    checkpoint = _module(
        'torchutil.checkpoint',
        load=_checkpoint_load,
        save=lambda *a, **k: None,
        latest_path=lambda *a, **k: None,
        best_path=lambda *a, **k: None)
    metrics = _module(
        'torchutil.metrics',
        Average=_Average,
        MeanStd=_MeanStd,
        PearsonCorrelation=_PearsonCorrelation)
    tensorboard = _module('torchutil.tensorboard', update=lambda *a, **k: None)
    download = _module('torchutil.download', file=lambda *a, **k: None)
    _module(
        'torchutil',
        notify=_notify,
        iterator=_iterator,
        multiprocess_iterator=lambda *a, **k: None,
        checkpoint=checkpoint,
        metrics=metrics,
        tensorboard=tensorboard,
        download=download)

    _module(
        'pypar',
        Alignment=Alignment,
        Word=Word,
        Phoneme=Phoneme,
        SILENCE=SILENCE)
    for name in ('pyfoal', 'penn', 'reseval', 'pycwt'):
        _module(name)
    pyplot = _module('matplotlib.pyplot')
    _module('matplotlib', pyplot=pyplot)

    def mel(sr, n_fft, n_mels):
        return emphases_oracle.mel_basis(sr, n_fft, n_mels)
    filters = _module('librosa.filters', mel=mel)
    _module('librosa', filters=filters)

    if 'huggingface_hub' not in sys.modules:
        try:
            importlib.import_module('huggingface_hub')
        except Exception:
            _module('huggingface_hub', hf_hub_download=None)


def import_reference(overrides=None):
    """Return a freshly imported reference `emphases` package"""
    for name in list(sys.modules):
        if name == 'emphases' or name.startswith('emphases.'):
            del sys.modules[name]
    install(overrides)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import emphases
    # mels.py caches the basis under the wrong attribute name (mels.py:96 vs
    # :103) so nothing stale can survive a re-import; the Hann window cache
    # lives on the fresh function object.
    return emphases


@contextlib.contextmanager
def reference(overrides=None):
    emphases = import_reference(overrides)
    try:
        yield emphases
    finally:
        for name in list(sys.modules):
            if name == 'emphases' or name.startswith('emphases.'):
                del sys.modules[name]
