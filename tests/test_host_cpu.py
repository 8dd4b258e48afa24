"""CPU tests of the host-side logic and of the C-ABI library's exports
(no compute calls without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

from emphases_b200 import _lib, engine
from golden_util import ROOT, times_list
from oracle import emphases_oracle as oracle


def test_library_exports_every_declared_symbol():
    from emphases_b200 import build
    build.build()
    header = open(os.path.join(ROOT, 'include', 'emphases_b200.h')).read()
    declared = set(re.findall(r'\b(emph_[a-z0-9_]+)\s*\(', header))
    assert len(declared) >= 10
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f'{name} declared but not exported'
    assert set(_lib.SIGNATURES) <= declared
    lib.emph_version.restype = ctypes.c_int
    assert lib.emph_version() >= 100


def test_mel_basis_matches_oracle(golden):
    basis = engine.mel_basis()
    np.testing.assert_array_equal(basis, oracle.mel_basis())
    np.testing.assert_array_equal(basis, golden('c1')['mel_basis'])
    ptr, col, val = engine.basis_to_csr(basis)
    assert ptr[-1] == 1001 and len(ptr) == 81
    dense = np.zeros_like(basis)
    for m in range(80):
        dense[m, col[ptr[m]:ptr[m + 1]]] = val[ptr[m]:ptr[m + 1]]
    np.testing.assert_array_equal(dense, basis)


@pytest.mark.parametrize('batch_size', [None, 300, 100, 1, 5000])
def test_chunker_bit_exact(golden, batch_size):
    """word bounds / chunk frames must match the reference bit-exactly"""
    data = golden('c1')
    times = np.asarray(data['times'])
    chunks = engine.chunk_words(times, 160000, batch_size)
    plan = oracle.chunk_plan(times_list(times), 160000, batch_size)
    assert len(chunks) == len(plan)
    for (w0, w1, start, length, bounds), expected in zip(chunks, plan):
        assert (w0, w1) == (expected['word_start'], expected['word_end'])
        assert start == expected['start_sample']
        assert length == expected['length']
        assert bounds.tolist() == [list(b) for b in expected['bounds']]
    tag = {None: 'full', 300: 'bs300', 100: 'bs100'}.get(batch_size)
    if tag:
        assert len(chunks) == int(data[f'{tag}.num_chunks'])
        for i, chunk in enumerate(chunks):
            np.testing.assert_array_equal(
                chunk[4].T[None], data[f'{tag}.{i}.bounds'])
            assert chunk[3] // 160 == int(data[f'{tag}.{i}.frames'])


def test_chunker_ragged_corpus_matches_oracle():
    for seed in range(40):
        times, audio = oracle.synthetic_utterance(1000 + seed)
        for batch_size in (None, 250):
            chunks = engine.chunk_words(
                np.asarray(times), audio.shape[-1], batch_size)
            plan = oracle.chunk_plan(times, audio.shape[-1], batch_size)
            assert len(chunks) == len(plan)
            for chunk, expected in zip(chunks, plan):
                assert chunk[2] == expected['start_sample']
                assert chunk[3] == expected['length']
                assert chunk[4].tolist() == [list(b) for b in expected['bounds']]


def test_chunker_edge_cases():
    # alignment longer than the audio: chunk clipped to the padded length
    times = np.array([[0.0, 0.5], [0.5, 1.2]])
    chunks = engine.chunk_words(times, 16000, None)
    plan = oracle.chunk_plan(times_list(times), 16000, None)
    assert [c[3] for c in chunks] == [p['length'] for p in plan]
    assert chunks[0][3] == 16000 + 864
    # chunk too short for the reflect pad is dropped
    times = np.array([[0.0, 0.02], [0.02, 1.0]])
    chunks = engine.chunk_words(times, 16000, 0)
    plan = oracle.chunk_plan(times_list(times), 16000, 0)
    assert len(chunks) == len(plan) == 1
    # empty alignment
    assert engine.chunk_words(np.zeros((0, 2)), 16000, None) == []


def test_plan_layout():
    utterances = []
    for seed in range(5):
        times, audio = oracle.synthetic_utterance(seed)
        utterances.append((np.asarray(times), audio.shape[-1]))
    plan = engine.make_plan(utterances)
    assert plan.n_seq == 5
    assert plan.row_start[0] == 1
    for u in range(1, 5):
        assert plan.row_start[u] == plan.row_start[u - 1] + plan.n_rows[u - 1] + 1
        assert plan.audio_off[u] % 4 == 0
    assert plan.total_rows == plan.row_start[-1] + plan.n_rows[-1] + 1
    assert (plan.word_seq >= 0).sum() == plan.n_words.sum()
    for u in range(5):
        s, n = plan.word_row_start[u], plan.n_words[u]
        assert (plan.word_seq[s:s + n] == u).all()
        assert plan.word_seq[s - 1] == -1 and plan.word_seq[s + n] == -1
        expected = oracle.word_bounds(
            [tuple(t) for t in utterances[u][0].tolist()])
        assert list(zip(plan.word_lo[s:s + n], plan.word_hi[s:s + n])) == expected


def test_validate_bounds_raises_like_reference(golden):
    data = golden('pool')
    bounds = np.stack([data['bounds'][0, 0], data['bounds'][0, 1]], axis=1)
    frames = np.full(len(bounds), 50)
    for method in ('max', 'center'):
        assert str(data[f'adversarial.{method}.error']) == 'IndexError'
        with pytest.raises(IndexError):
            engine.validate_bounds(bounds, frames, method)
    for method in ('sum', 'average'):
        engine.validate_bounds(bounds, frames, method)


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    with pytest.raises(_lib.EmphasesB200Error):
        engine.Engine('cpu')
