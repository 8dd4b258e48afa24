import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for path in (ROOT, os.path.join(ROOT, 'tests')):
    if path not in sys.path:
        sys.path.insert(0, path)

from golden_util import load_golden  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line(
        'markers', 'gpu: test needs a CUDA device (run on the B200 box)')


@pytest.fixture(scope='session')
def golden():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = load_golden(name)
        return cache[name]
    return get
