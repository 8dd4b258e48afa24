"""CPU tests of the training batch sampler / collation and of the oracle's
evaluation restatement, against goldens from the unmodified reference
(oracle/gen_golden.py: gen_sampler, gen_evaluate)."""
import numpy as np
import pytest
import torch

from oracle import emphases_oracle as oracle


@pytest.fixture
def emphases():
    import emphases_b200
    emphases_b200.reset_configuration()
    yield emphases_b200
    emphases_b200.reset_configuration()


def split(flat, sizes):
    edges = np.concatenate([[0], np.cumsum(sizes)])
    return [flat[a:b].tolist() for a, b in zip(edges[:-1], edges[1:])]


@pytest.mark.parametrize('name', ['a', 'b', 'c'])
@pytest.mark.parametrize('epoch', [0, 3])
def test_sampler_batches_equal_reference(emphases, golden, name, epoch):
    data = golden('sampler')
    lengths = data[f'{name}/lengths']
    max_frames = int(data[f'{name}/max_frames'])
    want = split(data[f'{name}/epoch{epoch}/flat'], data[f'{name}/epoch{epoch}/sizes'])

    assert [
        [int(i) for i in batch]
        for batch in oracle.epoch_batches(lengths, epoch, max_frames)] == want

    sampler = emphases.data.Sampler(emphases.data.LengthDataset(lengths), max_frames)
    sampler.set_epoch(epoch)
    got = [[int(i) for i in batch] for batch in sampler.batch()]
    assert got == want
    assert len(sampler) == len(want)
    # every item exactly once, every batch within the frame budget
    assert sorted(i for batch in got for i in batch) == list(range(len(lengths)))
    for batch in got:
        assert len(batch) == 1 or len(batch) * lengths[batch].max() <= max_frames
    # rank shards partition the epoch's batches
    shards = [sampler.shard(rank, 2) for rank in range(2)]
    assert sorted(map(tuple, shards[0] + shards[1])) == sorted(map(tuple, got))


def test_sampler_selection(emphases):
    dataset = emphases.data.LengthDataset([5, 9, 7, 3])
    assert isinstance(emphases.data.sampler(dataset, 'train'), emphases.data.Sampler)
    assert list(emphases.data.sampler(dataset, 'test')) == [[0], [1], [2], [3]]
    with pytest.raises(ValueError, match='Partition other is not defined'):
        emphases.data.sampler(dataset, 'other')


def test_collate_equals_reference(emphases, golden):
    data = golden('sampler')
    items = [
        (torch.from_numpy(data[f'collate/item{i}/features']),
         torch.from_numpy(data[f'collate/item{i}/scores']),
         torch.from_numpy(data[f'collate/item{i}/bounds']),
         None,
         torch.from_numpy(data[f'collate/item{i}/audio']),
         f'stem{i}')
        for i in range(3)]
    batch = emphases.data.collate(items)
    for key, value in zip(
        ('features', 'frame_lengths', 'word_bounds', 'word_lengths', 'scores',
         None, 'audio', None), batch
    ):
        if key is None:
            continue
        want = torch.from_numpy(data[f'collate/{key}'])
        assert value.dtype == want.dtype and value.shape == want.shape, key
        assert torch.equal(value, want), key
    assert batch[7] == ('stem0', 'stem1', 'stem2')


@pytest.mark.parametrize('loss_fn', ['bce', 'mse'])
def test_oracle_evaluate_matches_reference_metrics(golden, loss_fn):
    data = golden('evaluate')
    logits = [torch.from_numpy(data[f'logits{i}']).reshape(-1) for i in range(6)]
    targets = [torch.from_numpy(data[f'targets{i}']).reshape(-1) for i in range(6)]
    overall, granular = oracle.evaluate(logits, targets, loss_fn)
    keys = ('pearson_correlation', 'bce', 'mse')
    np.testing.assert_allclose(
        [overall[k] for k in keys], data[f'{loss_fn}/overall'], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(
        [[g[k] for k in keys] for g in granular], data[f'{loss_fn}/granular'],
        rtol=1e-5, atol=1e-6)
