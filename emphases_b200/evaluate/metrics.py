"""Evaluation metrics (emphases/evaluate/metrics.py:13-111) on the device.

The reference keeps three torchutil accumulators per file and per dataset and
feeds them one file at a time; here one kernel launch
(`emph_word_metric_sums`, csrc/eval.cu) produces fp64 per-file sums for the
whole dataset and these classes only combine them.  Same names, same
`update` / `reset` / `__call__` protocol, same result dictionary.

torchutil (third party, not vendored by the reference) defines the three
accumulators; their arithmetic as used here: Average = sum / count; MeanStd =
mean and SAMPLE standard deviation (n - 1); PearsonCorrelation =
sum((p - mean_p) (t - mean_t)) / (n std_p std_t).
"""
import math

import numpy as np
import torch

import emphases_b200 as emphases
from .. import _lib


def _loss_mode():
    if emphases.LOSS == 'bce':
        return 0
    if emphases.LOSS == 'mse':
        return 1
    raise ValueError(f'Loss {emphases.LOSS} is not recognized')


def word_sums(logits, targets, counts, passno, mean_p=0., mean_t=0.):
    """logits, targets: flat fp32 CUDA tensors of all words of all files in
    order; counts: words per file.  Returns an (n_files, 5) float64 array."""
    if not logits.is_cuda:
        raise _lib.EmphasesB200Error(
            'evaluation metrics run on the GPU; there is no CPU fallback')
    counts = np.asarray(counts, dtype=np.int64)
    starts = np.concatenate([[0], np.cumsum(counts[:-1])]) if len(counts) else counts
    if int(counts.sum()) != logits.numel() or logits.numel() != targets.numel():
        raise ValueError('word counts do not match the logits / targets')
    meta = torch.from_numpy(
        np.concatenate([starts, counts]).astype(np.int32)).to(logits.device)
    sums = torch.zeros((len(counts), 5), dtype=torch.float64, device=logits.device)
    logits = logits.contiguous().float()
    targets = targets.contiguous().float()
    _lib.call(
        'emph_word_metric_sums', _lib.ptr(logits), _lib.ptr(targets),
        _lib.ptr(meta), _lib.ptr(meta[len(counts):]), len(counts),
        _loss_mode(), passno, float(mean_p), float(mean_t), _lib.ptr(sums),
        _lib.stream_ptr())
    return sums.cpu().numpy()


###############################################################################
# Aggregate metric
###############################################################################


class Statistics:
    """metrics.py:103-111 (torchutil MeanStd): dataset mean / std"""

    def __init__(self):
        self.reset()

    def reset(self):
        self.count = 0
        self.total = 0.
        self.values = []

    def update(self, values, lengths=None):
        """values: (B, 1, Wmax) or flat tensor; lengths: (B,) valid words"""
        if lengths is not None:
            mask = emphases.model.mask_from_lengths(lengths.to(values.device))
            values = values[mask]
        self.values.append(values.detach().flatten().double().cpu())

    def __call__(self):
        values = torch.cat(self.values)
        mean = values.mean().item()
        std = math.sqrt(((values - mean) ** 2).sum().item() / (len(values) - 1))
        return mean, std


class Metrics:
    """metrics.py:13-52.  `update(logits, targets, word_lengths)` takes the
    reference's padded (B, 1, Wmax) tensors; `update_packed` takes flat
    per-word tensors plus words-per-file and also returns per-file results."""

    def __init__(self, predicted_stats, target_stats):
        self.mean, self.std = predicted_stats()
        self.target_mean, self.target_std = target_stats()
        self.reset()

    def reset(self):
        self.sums = np.zeros(5)
        self.count = 0

    def update(self, logits, targets, word_lengths):
        mask = emphases.model.mask_from_lengths(word_lengths.to(logits.device))
        logits = logits.detach()[mask]
        targets = targets.to(logits.device)[mask]
        self.update_packed(logits, targets, [logits.numel()])

    def update_packed(self, logits, targets, counts):
        sums = word_sums(
            logits, targets, counts, 1, self.mean, self.target_mean)
        self.sums += sums.sum(0)
        self.count += int(np.sum(counts))
        return [self._result(s, c) for s, c in zip(sums, counts)]

    def _result(self, sums, count):
        return {
            'pearson_correlation': float(
                1. / count * (sums[2] / (self.std * self.target_std))),
            'bce': float(sums[3] / count),
            'mse': float(sums[4] / count)}

    def __call__(self):
        return self._result(self.sums, self.count)
