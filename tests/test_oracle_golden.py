"""Pin the CPU oracle (oracle/emphases_oracle.py) against golden vectors
produced by the unmodified reference (oracle/gen_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import emphases_oracle as oracle
from golden_util import state_from_golden, times_list

METHODS = ['average', 'max', 'sum', 'center']
LOCATIONS = ['input', 'intermediate', 'inference', 'loss']


def test_known_answer_bounds(golden):
    """SURVEY.md section 8c known answer: word bounds of the C1 example"""
    data = golden('c1')
    bounds = data['full.0.bounds'][0]
    assert bounds[0, :6].tolist() == [0, 29, 198, 279, 284, 305]
    assert bounds[:, -3:].T.tolist() == [[757, 799], [799, 937], [937, 1000]]
    assert oracle.word_bounds(times_list(data['times'])) == [
        tuple(pair) for pair in bounds.T.tolist()]


def test_known_answer_scores(golden):
    data = golden('c1')
    expected = [3.9732e-02, 7.6509e-01, 3.9910e-01, 4.6730e-05, 4.3213e-02,
                2.1440e-02]
    np.testing.assert_allclose(
        data['full.scores'][0, :6], expected, rtol=2e-4)


@pytest.mark.parametrize('batch_size', [None, 300, 100])
def test_chunker_and_features(golden, batch_size):
    """emphases.preprocess: chunk count, bit-exact bounds, frame counts,
    log-mel features"""
    data = golden('c1')
    tag = 'full' if batch_size is None else f'bs{batch_size}'
    times = times_list(data['times'])
    audio = torch.from_numpy(data['audio'])
    plan = oracle.chunk_plan(times, audio.shape[-1], batch_size)
    assert len(plan) == int(data[f'{tag}.num_chunks'])
    chunks = list(oracle.preprocess(
        times, audio, batch_size, data['mel_basis']))
    for i, (features, bounds) in enumerate(chunks):
        assert np.array_equal(bounds.numpy(), data[f'{tag}.{i}.bounds'])
        assert features.shape[-1] == int(data[f'{tag}.{i}.frames'])
        assert plan[i]['frames'] == features.shape[-1]
        if f'{tag}.{i}.features' in data:
            np.testing.assert_allclose(
                features.numpy(), data[f'{tag}.{i}.features'],
                rtol=0, atol=1e-6)


@pytest.mark.parametrize('batch_size', [None, 300, 100])
def test_end_to_end_scores(golden, batch_size):
    data = golden('c1')
    tag = 'full' if batch_size is None else f'bs{batch_size}'
    state = state_from_golden(data)
    scores = oracle.from_alignment_and_audio(
        times_list(data['times']),
        torch.from_numpy(data['audio']),
        state,
        batch_size=batch_size,
        basis=data['mel_basis'])
    np.testing.assert_allclose(
        scores.numpy(), data[f'{tag}.scores'], rtol=0, atol=2e-7)


def test_intermediates(golden):
    data = golden('c1')
    state = state_from_golden(data)
    features = torch.from_numpy(data['full.0.features'])
    bounds = torch.from_numpy(data['full.0.bounds'])
    with torch.no_grad():
        _, inter = oracle.model_forward(
            state, features, torch.tensor([1000]), bounds,
            torch.tensor([25]), return_intermediates=True)
    np.testing.assert_allclose(
        inter['frame_embeddings'].numpy(), data['full.frame_embeddings'],
        rtol=0, atol=1e-6)
    np.testing.assert_allclose(
        inter['word_embeddings'].numpy(), data['full.word_embeddings'],
        rtol=0, atol=1e-4)


def test_autocast_baseline_close(golden):
    """The timed CPU baseline mode (reference's own bf16 autocast)"""
    data = golden('c1')
    state = state_from_golden(data)
    scores = oracle.from_alignment_and_audio(
        times_list(data['times']), torch.from_numpy(data['audio']), state,
        basis=data['mel_basis'], autocast=True)
    assert str(data['full.scores_autocast_dtype']) == 'torch.bfloat16'
    assert scores.dtype == torch.bfloat16
    np.testing.assert_allclose(
        scores.float().numpy(), data['full.scores_autocast'], atol=4e-3)


@pytest.mark.parametrize('location', LOCATIONS)
@pytest.mark.parametrize('method', METHODS)
def test_sweep(golden, location, method):
    data = golden('sweep')
    state = state_from_golden(data)
    config = {'DOWNSAMPLE_LOCATION': location, 'DOWNSAMPLE_METHOD': method}
    tag = f'{location}.{method}'
    with torch.no_grad():
        features = torch.from_numpy(data['b1.features'])
        bounds = torch.from_numpy(data['b1.bounds'])
        logits = oracle.model_forward(
            state, features, torch.tensor([features.shape[-1]]), bounds,
            torch.tensor([bounds.shape[-1]]), config)
        np.testing.assert_allclose(
            logits.numpy(), data[f'{tag}.b1.logits'], rtol=1e-5, atol=1e-5)
        batch = [torch.from_numpy(data[f'b2.{name}']) for name in (
            'features', 'frame_lengths', 'bounds', 'word_lengths')]
        logits = oracle.model_forward(state, *batch, config)
        np.testing.assert_allclose(
            logits.numpy(), data[f'{tag}.b2.logits'], rtol=1e-5, atol=1e-5)
        if location == 'inference':
            frame_logits = oracle.model_forward(
                state, *batch, config, training=True)
            np.testing.assert_allclose(
                frame_logits.numpy(), data[f'{tag}.b2.frame_logits'],
                rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize('method', METHODS)
def test_downsample(golden, method):
    data = golden('pool')
    xs = torch.from_numpy(data['xs'])
    result = oracle.downsample(
        xs, torch.from_numpy(data['clean_bounds']),
        torch.from_numpy(data['clean_lengths']), method)
    np.testing.assert_allclose(
        result.numpy(), data[f'clean.{method}'], rtol=0, atol=0)
    error = str(data[f'adversarial.{method}.error'])
    bounds = torch.from_numpy(data['bounds'])
    lengths = torch.from_numpy(data['lengths'])
    if error:
        with pytest.raises(IndexError):
            oracle.downsample(xs, bounds, lengths, method)
    else:
        result = oracle.downsample(xs, bounds, lengths, method)
        np.testing.assert_array_equal(
            result.numpy(), data[f'adversarial.{method}'])


def test_segment(golden):
    data = golden('pool')
    segments, bounds, lengths = oracle.segment(
        torch.from_numpy(data['xs']),
        torch.from_numpy(data['clean_bounds']),
        torch.from_numpy(data['clean_lengths']))
    np.testing.assert_array_equal(segments.numpy(), data['segment.segments'])
    np.testing.assert_array_equal(bounds.numpy(), data['segment.bounds'])
    np.testing.assert_array_equal(lengths.numpy(), data['segment.lengths'])


def test_transformer(golden):
    data = golden('transformer')
    state = state_from_golden(data)
    encoding = oracle.positional_encoding(80)
    state['frame_encoder.position.encoding'] = encoding
    state['word_decoder.position.encoding'] = encoding
    config = {'ARCHITECTURE': 'transformer'}
    with torch.no_grad():
        features = torch.from_numpy(data['b1.features'])
        bounds = torch.from_numpy(data['b1.bounds'])
        logits, inter = oracle.model_forward(
            state, features, torch.tensor([features.shape[-1]]), bounds,
            torch.tensor([bounds.shape[-1]]), config,
            return_intermediates=True)
        np.testing.assert_allclose(
            inter['frame_embeddings'].numpy(), data['b1.frame_embeddings'],
            rtol=0, atol=2e-5)
        np.testing.assert_allclose(
            logits.numpy(), data['b1.logits'], rtol=0, atol=2e-5)
        batch = [torch.from_numpy(data[f'b2.{name}']) for name in (
            'features', 'frame_lengths', 'bounds', 'word_lengths')]
        logits = oracle.model_forward(state, *batch, config)
        # padded word slots of the shorter item hold garbage-in-reference
        # values that depend on masked softmax rows; compare valid slots
        for i, words in enumerate(batch[3].tolist()):
            np.testing.assert_allclose(
                logits.numpy()[i, :, :words],
                data['b2.logits'][i, :, :words], rtol=0, atol=2e-5)


def test_transformer_input_location(golden):
    """Transformer variant with the word segments as attention sequences
    (DOWNSAMPLE_LOCATION='input', emphases/model/core.py:41-87)"""
    data = golden('transformer_input')
    state = state_from_golden(data)
    encoding = oracle.positional_encoding(80)
    state['frame_encoder.position.encoding'] = encoding
    state['word_decoder.position.encoding'] = encoding
    config = {'ARCHITECTURE': 'transformer', 'DOWNSAMPLE_LOCATION': 'input'}
    with torch.no_grad():
        features = torch.from_numpy(data['b1.features'])
        bounds = torch.from_numpy(data['b1.bounds'])
        logits = oracle.model_forward(
            state, features, torch.tensor([features.shape[-1]]), bounds,
            torch.tensor([bounds.shape[-1]]), config)
        np.testing.assert_allclose(
            logits.numpy(), data['b1.logits'], rtol=0, atol=2e-5)
        batch = [torch.from_numpy(data[f'b2.{name}']) for name in (
            'features', 'frame_lengths', 'bounds', 'word_lengths')]
        logits = oracle.model_forward(state, *batch, config)
        for i, words in enumerate(batch[3].tolist()):
            np.testing.assert_allclose(
                logits.numpy()[i, :, :words],
                data['b2.logits'][i, :, :words], rtol=0, atol=2e-5)


@pytest.mark.parametrize('loss_fn', ['bce', 'mse'])
def test_loss(golden, loss_fn):
    data = golden('loss')
    value = oracle.loss(
        torch.from_numpy(data['scores']), torch.from_numpy(data['targets']),
        torch.from_numpy(data['word_lengths']), loss_fn)
    np.testing.assert_allclose(value.numpy(), data[loss_fn], rtol=1e-6)


@pytest.mark.parametrize('method', ['linear', 'nearest'])
def test_upsample_and_frame_loss(golden, method):
    data = golden('upsample')
    bounds = torch.from_numpy(data['bounds'])
    word_lengths = torch.from_numpy(data['word_lengths'])
    frame_lengths = torch.from_numpy(data['frame_lengths'])
    for name in ('xs', 'wide'):
        result = oracle.upsample(
            torch.from_numpy(data[name]), bounds, word_lengths, frame_lengths,
            method)
        np.testing.assert_allclose(
            result.numpy(), data[f'{method}.{name}'], rtol=1e-6, atol=1e-6)
    for loss_fn in ('bce', 'mse'):
        value = oracle.loss(
            torch.from_numpy(data['scores']), torch.from_numpy(data['xs']),
            word_lengths, loss_fn, frame_lengths, bounds, method)
        np.testing.assert_allclose(
            value.numpy(), data[f'{method}.loss.{loss_fn}'], rtol=1e-6)
