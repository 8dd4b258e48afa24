"""Helpers shared by the test modules (golden loading)."""
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, f'{name}.npz')))


def state_from_golden(data, prefix='state.'):
    return {
        key[len(prefix):]: torch.from_numpy(value)
        for key, value in data.items() if key.startswith(prefix)}


def times_list(array):
    return [(float(a), float(b)) for a, b in array]
