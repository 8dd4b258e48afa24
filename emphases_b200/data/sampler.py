"""Length-bucketed batch sampler feeding the data-parallel training step.

Mirrors emphases/data/sampler.py:11-87 and Dataset.buckets
(emphases/data/dataset.py:97-117): same bucket boundaries, same
`torch.Generator(RANDOM_SEED + epoch)` shuffles, same greedy variable batch
size under `MAX_TRAINING_FRAMES`, so a given (lengths, epoch) yields the same
batches as the reference.  `shard` is the extension: rank r of a
`world_size`-process job takes batches r, r + world_size, ... of the epoch's
batch list (every rank derives the identical list from the shared seed, so no
communication is needed).
"""
import numpy as np
import torch

import emphases_b200 as emphases


def buckets(lengths, count=None):
    """dataset.py:97-117: indices in order of length, cut into `BUCKETS`
    equal parts, the remainder folded into the last one.  Returns a list of
    (n, 2) int arrays of (index, length)."""
    count = emphases.BUCKETS if count is None else count
    lengths = np.asarray(lengths)
    total = len(lengths)
    size = total // count
    if size == 0:
        raise ValueError(
            f'{total} items cannot fill {count} buckets '
            '(range() arg 3 must not be zero, dataset.py:107)')
    order = np.argsort(lengths)
    ordered = np.sort(lengths)
    parts = [
        np.stack((order[i:i + size], ordered[i:i + size])).T
        for i in range(0, total, size)]
    if len(parts) == count + 1:
        residual = parts.pop()
        parts[-1] = np.concatenate((parts[-1], residual), axis=0)
    return parts


class Sampler:
    """sampler.py:33-87"""

    def __init__(self, dataset, max_frames=None):
        self.max_frames = (
            emphases.MAX_TRAINING_FRAMES if max_frames is None else max_frames)
        self.epoch = 0
        self.length = len(dataset)
        self.buckets = dataset.buckets()

    def __iter__(self):
        return iter(self.batch())

    def __len__(self):
        return len(self.batch())

    def batch(self):
        """Batch indices for one epoch"""
        generator = torch.Generator()
        generator.manual_seed(emphases.RANDOM_SEED + self.epoch)
        batches = []
        for bucket in self.buckets:
            order = torch.randperm(len(bucket), generator=generator).tolist()
            batch, longest = [], 0
            for index, length in bucket[order]:
                longest = max(longest, length)
                if batch and (len(batch) + 1) * longest > self.max_frames:
                    batches.append(batch)
                    longest = length
                    batch = [index]
                else:
                    batch.append(index)
            if batch:
                batches.append(batch)
        order = torch.randperm(len(batches), generator=generator).tolist()
        return [batches[i] for i in order]

    def set_epoch(self, epoch):
        self.epoch = epoch

    def shard(self, rank, world_size):
        """This rank's batches of the current epoch (extension)"""
        return self.batch()[rank::world_size]


class LengthDataset:
    """The two members `Sampler` needs, for corpora that are not on disk in
    the reference's cache layout (synthetic benchmarks, tests)"""

    def __init__(self, lengths):
        self.lengths = list(lengths)

    def __len__(self):
        return len(self.lengths)

    def buckets(self):
        return buckets(self.lengths)


def sampler(dataset, partition):
    """sampler.py:11-25"""
    if partition in ['train', 'valid']:
        return Sampler(dataset)
    elif partition.startswith('test'):
        return torch.utils.data.BatchSampler(
            torch.utils.data.SequentialSampler(dataset), 1, False)
    raise ValueError(f'Partition {partition} is not defined')
