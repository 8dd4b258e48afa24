"""GPU parity tests of the evaluation caller (emphases_b200.evaluate) against
the metrics of the unmodified reference (tests/golden/evaluate.npz) and the
oracle."""
import json

import numpy as np
import pytest
import torch

from golden_util import state_from_golden
from oracle import emphases_oracle as oracle

pytestmark = pytest.mark.gpu

KEYS = ('pearson_correlation', 'bce', 'mse')


@pytest.fixture
def emphases():
    import emphases_b200
    emphases_b200.reset_configuration()
    yield emphases_b200
    emphases_b200.reset_configuration()


@pytest.mark.parametrize('loss_fn', ['bce', 'mse'])
def test_metrics_equal_reference(emphases, golden, loss_fn):
    """Statistics / Metrics fed like evaluate/core.py:27-110 feeds them"""
    emphases.configure(LOSS=loss_fn)
    data = golden('evaluate')
    logits = [torch.from_numpy(data[f'logits{i}']).cuda() for i in range(6)]
    targets = [torch.from_numpy(data[f'targets{i}']) for i in range(6)]
    target_stats = emphases.evaluate.metrics.Statistics()
    predicted_stats = emphases.evaluate.metrics.Statistics()
    for x, t in zip(logits, targets):
        lengths = torch.tensor([x.shape[-1]])
        target_stats.update(t, lengths)
        predicted_stats.update(emphases.postprocess(x[0])[None], lengths)
    np.testing.assert_allclose(
        [*predicted_stats(), *target_stats()], data[f'{loss_fn}/stats'], rtol=1e-6)
    file_metrics = emphases.evaluate.Metrics(predicted_stats, target_stats)
    dataset_metrics = emphases.evaluate.Metrics(predicted_stats, target_stats)
    for index, (x, t) in enumerate(zip(logits, targets)):
        lengths = torch.tensor([x.shape[-1]])
        file_metrics.reset()
        file_metrics.update(x, t, lengths)
        dataset_metrics.update(x, t, lengths)
        got = file_metrics()
        np.testing.assert_allclose(
            [got[k] for k in KEYS], data[f'{loss_fn}/granular'][index],
            rtol=1e-5, atol=1e-6)
    got = dataset_metrics()
    np.testing.assert_allclose(
        [got[k] for k in KEYS], data[f'{loss_fn}/overall'], rtol=1e-5, atol=1e-6)


def test_padded_batch_update(emphases, golden):
    """(B, 1, Wmax) batches with a length mask, metrics.py:31-46"""
    data = golden('evaluate')
    sizes = [data[f'logits{i}'].shape[-1] for i in range(6)]
    logits = torch.zeros(6, 1, max(sizes))
    targets = torch.zeros(6, 1, max(sizes))
    for i, n in enumerate(sizes):
        logits[i, :, :n] = torch.from_numpy(data[f'logits{i}'])[0]
        targets[i, :, :n] = torch.from_numpy(data[f'targets{i}'])[0]
    lengths = torch.tensor(sizes)
    stats = [emphases.evaluate.metrics.Statistics() for _ in range(2)]
    stats[0].update(emphases.postprocess(logits), lengths)
    stats[1].update(targets, lengths)
    metrics = emphases.evaluate.Metrics(*stats)
    metrics.update(logits.cuda(), targets, lengths)
    got = metrics()
    np.testing.assert_allclose(
        [got[k] for k in KEYS], data['bce/overall'], rtol=1e-5, atol=1e-6)


def make_dataset(emphases, count, seed):
    generator = torch.Generator().manual_seed(seed)
    items = []
    for index in range(count):
        times, audio = oracle.synthetic_utterance(seed + index)
        target = torch.rand(len(times), generator=generator)
        items.append((times, audio, target))
    return items


def oracle_logits(times, audio, state):
    pieces = []
    for features, bounds in oracle.preprocess(times, audio):
        pieces.append(oracle.model_forward(
            state, features, torch.tensor([features.shape[-1]]), bounds,
            torch.tensor([bounds.shape[-1]]))[0, 0])
    return torch.cat(pieces)


def test_corpus_matches_oracle(emphases, golden, tmp_path):
    state = state_from_golden(golden('c1'))
    checkpoint = tmp_path / 'checkpoint.pt'
    torch.save({'model': state}, checkpoint)
    items = make_dataset(emphases, 7, 900)
    want_overall, want_files = oracle.evaluate(
        [oracle_logits(times, audio, state) for times, audio, _ in items],
        [target for _, _, target in items])
    overall, granular = emphases.evaluate.corpus(
        [emphases.Alignment.from_times(times) for times, _, _ in items],
        [audio for _, audio, _ in items],
        [target[None] for _, _, target in items],
        checkpoint=checkpoint, gpu=0, stems=[f's{i}' for i in range(7)])
    np.testing.assert_allclose(
        [overall[k] for k in KEYS], [want_overall[k] for k in KEYS],
        rtol=1e-4, atol=1e-5)
    assert list(granular) == [f's{i}' for i in range(7)]
    for got, want in zip(granular.values(), want_files):
        np.testing.assert_allclose(
            [got[k] for k in KEYS], [want[k] for k in KEYS], rtol=1e-4, atol=1e-5)
    with pytest.raises(ValueError, match='targets for'):
        emphases.evaluate.corpus(
            [emphases.Alignment.from_times(items[0][0])], [items[0][1]],
            [items[0][2][:-1]], checkpoint=checkpoint, gpu=0)


def test_datasets_from_cache_layout(emphases, golden, tmp_path):
    """evaluate.datasets on the reference's on-disk layout"""
    state = state_from_golden(golden('c1'))
    checkpoint = tmp_path / 'checkpoint.pt'
    torch.save({'model': state}, checkpoint)
    cache = tmp_path / 'cache' / 'toy'
    for name in ('alignment', 'audio', 'scores'):
        (cache / name).mkdir(parents=True)
    (tmp_path / 'partitions').mkdir()
    items = make_dataset(emphases, 4, 950)
    stems = [f'file{i}' for i in range(4)]
    logits, targets = [], []
    for stem, (times, audio, target) in zip(stems, items):
        emphases.Alignment.from_times(times).save(cache / 'alignment' / f'{stem}.TextGrid')
        emphases.load.save_wav(cache / 'audio' / f'{stem}.wav', audio)
        torch.save(target[None], cache / 'scores' / f'{stem}.pt')
        loaded = emphases.load.audio(cache / 'audio' / f'{stem}.wav')
        logits.append(oracle_logits(times, loaded, state))
        targets.append(target)
    with open(tmp_path / 'partitions' / 'toy.json', 'w') as file:
        json.dump({'train': [], 'valid': [], 'test': stems}, file)
    emphases.configure(
        CACHE_DIR=tmp_path / 'cache', PARTITION_DIR=tmp_path / 'partitions',
        EVAL_DIR=tmp_path / 'eval')
    overall, granular = emphases.evaluate.datasets(['toy'], checkpoint, gpu=0)
    want_overall, want_files = oracle.evaluate(logits, targets)
    written = json.load(open(tmp_path / 'eval' / 'emphases' / 'overall.json'))
    assert written == overall
    np.testing.assert_allclose(
        [overall['toy'][k] for k in KEYS], [want_overall[k] for k in KEYS],
        rtol=1e-4, atol=1e-5)
    files = json.load(open(tmp_path / 'eval' / 'emphases' / 'granular.json'))
    assert list(files) == [f'toy/{stem}' for stem in stems]
    for stem, want in zip(stems, want_files):
        np.testing.assert_allclose(
            [files[f'toy/{stem}'][k] for k in KEYS], [want[k] for k in KEYS],
            rtol=1e-4, atol=1e-5)
