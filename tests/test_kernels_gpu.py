"""GPU parity tests: each sm_100a kernel, called through the C ABI, against
the CPU oracle and the golden vectors of the unmodified reference."""
import numpy as np
import pytest
import torch

from golden_util import state_from_golden, times_list
from oracle import emphases_oracle as oracle

pytestmark = pytest.mark.gpu

METHODS = ['average', 'max', 'sum', 'center']


@pytest.fixture(scope='module')
def eng():
    from emphases_b200 import engine
    return engine.Engine('cuda:0')


def default_weights(state, has_decoder=True):
    from emphases_b200 import engine
    return engine.pack_weights(
        state, torch.device('cuda:0'), layers=6, activation='ReLU',
        dropout=None, has_decoder=has_decoder)


def make_rows(lengths):
    """Packed layout helpers on device"""
    from emphases_b200 import engine
    starts, total = engine.packed_starts(lengths)
    row_start = torch.tensor(starts, dtype=torch.int32, device='cuda:0')
    n_rows = torch.tensor(lengths, dtype=torch.int32, device='cuda:0')
    return row_start, n_rows, total


def test_row_index(eng):
    lengths = [5, 1, 7, 3]
    row_start, n_rows, total = make_rows(lengths)
    row_seq = eng.row_index(row_start, n_rows, len(lengths), total).cpu().numpy()
    expected = np.full(total, -1)
    cursor = 1
    for u, n in enumerate(lengths):
        expected[cursor:cursor + n] = u
        cursor += n + 1
    np.testing.assert_array_equal(row_seq, expected)


@pytest.mark.parametrize('tag,count', [('full', 1), ('bs300', 3)])
@pytest.mark.parametrize('dtype', ['f32', 'i16'])
def test_logmel_golden(eng, golden, tag, count, dtype):
    """log-mel of every chunk of the C1 example vs the reference's features"""
    from emphases_b200 import engine
    data = golden('c1')
    times = np.asarray(data['times'])
    audio = torch.from_numpy(data['audio'])
    if dtype == 'i16':
        pcm = (audio * 32768.).round().clamp(-32768, 32767).to(torch.int16)
        audio = pcm.float() / 32768.
        device_audio = pcm[0].cuda()
    else:
        device_audio = audio[0].cuda()
    batch_size = None if tag == 'full' else 300
    plan = engine.make_plan([(times, audio.shape[-1])], batch_size)
    assert plan.n_seq == count
    views = eng.upload_plan(plan)
    row_seq = eng.row_index(
        views['row_start'], views['n_rows'], plan.n_seq, plan.total_rows)
    out = eng.logmel(device_audio, views, plan, row_seq).cpu()
    expected_chunks = list(oracle.preprocess(
        times_list(times), audio, batch_size, data['mel_basis']))
    for u in range(count):
        rows = out[plan.row_start[u]:plan.row_start[u] + plan.n_rows[u]]
        expected = expected_chunks[u][0][0].T
        assert rows.shape == expected.shape
        error = (rows - expected).abs().max().item()
        assert error < 2e-5, f'chunk {u}: log-mel max-abs {error}'
        if dtype == 'f32':
            np.testing.assert_allclose(
                rows.numpy(), data[f'{tag}.{u}.features'][0].T, atol=2e-5)
        # separator rows are zero
        assert out[plan.row_start[u] - 1].abs().max() == 0
    assert out[-1].abs().max() == 0


def test_logmel_ragged_edges(eng):
    """clipped chunks (alignment past the audio end), odd lengths, tiny chunks"""
    from emphases_b200 import engine
    generator = torch.Generator().manual_seed(3)
    utterances, audios = [], []
    for samples, times in [
        (16000, [[0.0, 0.5], [0.5, 1.2]]),          # alignment past the end
        (7777, [[0.0, 0.2], [0.2, 0.45]]),          # odd length
        (1600, [[0.03, 0.09]]),                     # tiny: 960 samples
        (48000, [[0.5, 1.0], [1.0, 2.9]]),          # first word starts late
    ]:
        audio = 0.1 * torch.randn(1, samples, generator=generator)
        utterances.append((np.asarray(times), samples))
        audios.append(audio)
    plan = engine.make_plan(utterances)
    packed = torch.zeros(plan.audio_samples)
    for offset, audio in zip(plan.audio_offsets, audios):
        packed[offset:offset + audio.shape[-1]] = audio[0]
    views = eng.upload_plan(plan)
    row_seq = eng.row_index(
        views['row_start'], views['n_rows'], plan.n_seq, plan.total_rows)
    out = eng.logmel(packed.cuda(), views, plan, row_seq).cpu()
    for u in range(plan.n_seq):
        index = int(plan.utterance[u])
        expected = list(oracle.preprocess(
            times_list(utterances[index][0]), audios[index]))[0][0][0].T
        rows = out[plan.row_start[u]:plan.row_start[u] + plan.n_rows[u]]
        assert rows.shape == expected.shape
        assert (rows - expected).abs().max().item() < 2e-5


def _logmel_corpus(seed=11):
    """A few utterances long enough for bulk-staged interior tiles, plus short ones"""
    from emphases_b200 import engine
    generator = torch.Generator().manual_seed(seed)
    utterances, audios = [], []
    for samples, words in [(52000, 7), (9000, 2), (160000, 20), (31111, 4), (16 * 160 + 1000, 2)]:
        audio = (0.1 * torch.randn(1, samples, generator=generator)).clamp(-1, 1)
        duration = samples / 16000.
        cuts = torch.sort(torch.rand(words - 1, generator=generator) * duration).values
        edges = np.concatenate([[0.], cuts.numpy(), [duration]])
        utterances.append((np.stack([edges[:-1], edges[1:]], axis=1), samples))
        audios.append(audio)
    plan = engine.make_plan(utterances)
    packed = torch.zeros(plan.audio_samples)
    for offset, audio in zip(plan.audio_offsets, audios):
        packed[offset:offset + audio.shape[-1]] = audio[0]
    return utterances, audios, plan, packed


def _check_logmel(out, plan, utterances, audios, basis=None, tolerance=2e-5):
    for u in range(plan.n_seq):
        index = int(plan.utterance[u])
        expected = list(oracle.preprocess(
            times_list(utterances[index][0]), audios[index], None, basis))[0][0][0].T
        rows = out[plan.row_start[u]:plan.row_start[u] + plan.n_rows[u]]
        assert rows.shape == expected.shape
        assert (rows - expected).abs().max().item() < tolerance
        assert out[plan.row_start[u] - 1].abs().max() == 0      # separator row


@pytest.mark.parametrize('dtype', ['f32', 'i16'])
def test_logmel_staged_tiles_match_per_frame_path(eng, dtype):
    """Interior tiles are fetched by cp.async.bulk when the span is 16-byte
    aligned; a buffer shifted by 4 samples (8 bytes of int16) sends every tile
    down the per-frame vector-load path.  Both must agree bit for bit and
    match the oracle."""
    utterances, audios, plan, packed = _logmel_corpus()
    if dtype == 'i16':
        pcm = (packed * 32768.).round().clamp(-32768, 32767).to(torch.int16)
        audios = [
            (a * 32768.).round().clamp(-32768, 32767).to(torch.int16).float() / 32768.
            for a in audios]
        packed = pcm
    views = eng.upload_plan(plan)
    row_seq = eng.row_index(
        views['row_start'], views['n_rows'], plan.n_seq, plan.total_rows)
    staged = eng.logmel(packed.cuda(), views, plan, row_seq).cpu().clone()
    _check_logmel(staged, plan, utterances, audios)
    # 32-byte aligned allocation + 4 samples: still 16-byte aligned for fp32,
    # only 8-byte aligned for int16
    shifted = torch.zeros(packed.numel() + 8, dtype=packed.dtype, device='cuda:0')
    shifted[4:4 + packed.numel()] = packed.cuda()
    other = eng.logmel(shifted[4:], views, plan, row_seq).cpu()
    assert torch.equal(staged, other)


@pytest.mark.parametrize('kind', ['narrow33', 'scattered'])
def test_logmel_other_bases(kind):
    """The banded mel table is built from whatever CSR basis is passed: an odd
    number of triangular filters, and a basis too scattered for the table
    (the kernel then reads the CSR entries from global memory)."""
    from emphases_b200 import engine
    generator = np.random.default_rng(5)
    if kind == 'narrow33':
        basis = engine.mel_basis(n_mels=33)
    else:
        basis = np.zeros((80, 513), dtype=np.float32)
        for m in range(80):
            cols = generator.choice(513, size=24, replace=False)
            basis[m, cols] = generator.uniform(0.001, 0.02, size=24).astype(np.float32)
    local = engine.Engine('cuda:0', n_mels=basis.shape[0])
    ptr, col, val = engine.basis_to_csr(basis)
    local.mel_ptr = torch.from_numpy(ptr).cuda()
    local.mel_col = torch.from_numpy(col).cuda()
    local.mel_val = torch.from_numpy(val).cuda()
    utterances, audios, plan, packed = _logmel_corpus(seed=12)
    views = local.upload_plan(plan)
    row_seq = local.row_index(
        views['row_start'], views['n_rows'], plan.n_seq, plan.total_rows)
    out = local.logmel(packed.cuda(), views, plan, row_seq).cpu()
    assert out.shape[1] == basis.shape[0]
    _check_logmel(out, plan, utterances, audios, basis=basis)


def oracle_conv_rows(state, prefix_layers, x_rows, lengths):
    """Apply convs per sequence with the oracle, return packed rows"""
    from emphases_b200 import engine
    starts, total = engine.packed_starts(lengths)
    out = torch.zeros(total, x_rows.shape[1])
    for start, n in zip(starts, lengths):
        x = x_rows[start:start + n].T[None]
        with torch.no_grad():
            for weight, bias, relu in prefix_layers:
                x = oracle._conv(x, weight, bias)
                if relu:
                    x = torch.relu(x)
        out[start:start + n] = x[0].T
    return out


@pytest.mark.parametrize('which', ['frame', 'word'])
def test_conv_stack_f32(eng, golden, which):
    from emphases_b200 import _lib
    data = golden('c1')
    state = state_from_golden(data)
    weights = default_weights(state)
    generator = torch.Generator().manual_seed(4)
    lengths = [300, 1, 2, 131, 57, 640]
    row_start, n_rows, total = make_rows(lengths)
    row_seq = eng.row_index(row_start, n_rows, len(lengths), total)
    x = torch.randn(total, 80, generator=generator)
    x[(row_seq < 0).cpu()] = 0
    if which == 'frame':
        layers = [(state['input_layer.weight'], state['input_layer.bias'], False)]
        layers += [
            (state[f'frame_encoder.{2 * i}.weight'],
             state[f'frame_encoder.{2 * i}.bias'], True) for i in range(6)]
        stack = weights.frame
    else:
        layers = [
            (state[f'word_decoder.{2 * i}.weight'],
             state[f'word_decoder.{2 * i}.bias'], True) for i in range(6)]
        stack = weights.word
    y = eng.conv_stack(x.cuda(), row_seq, stack, _lib.PREC_FP32).cpu()
    expected = oracle_conv_rows(state, layers, x, lengths)
    scale = expected.abs().max().item()
    error = (y - expected).abs().max().item()
    assert error < 2e-6 * max(scale, 1.0), (error, scale)
    assert y[(row_seq < 0).cpu()].abs().max() == 0


@pytest.mark.parametrize('method', METHODS)
def test_pool_clean(eng, golden, method):
    data = golden('pool')
    xs = torch.from_numpy(data['xs'])                     # (2, 80, 50)
    bounds = data['clean_bounds']
    lengths = data['clean_lengths']
    pooled = run_pool(eng, xs, bounds, lengths, method)
    np.testing.assert_allclose(
        pooled, data[f'clean.{method}'], rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize('method', ['average', 'sum'])
def test_pool_adversarial(eng, golden, method):
    """zero-length words, single-frame words, bounds past T, late first word"""
    data = golden('pool')
    xs = torch.from_numpy(data['xs'])
    pooled = run_pool(eng, xs, data['bounds'], data['lengths'], method)
    expected = data[f'adversarial.{method}']
    np.testing.assert_array_equal(np.isnan(pooled), np.isnan(expected))
    np.testing.assert_allclose(
        np.nan_to_num(pooled), np.nan_to_num(expected), rtol=1e-6, atol=1e-6)


def run_pool(eng, xs, bounds, lengths, method):
    """(B, C, T) + padded bounds -> (B, C, Wmax) through emph_pool_words"""
    B, C, T = xs.shape
    wmax = bounds.shape[2]
    row_start, n_rows, total = make_rows([T] * B)
    word_row_start, n_words, total_words = make_rows([wmax] * B)
    rows = torch.zeros(total, C)
    for b in range(B):
        rows[int(row_start[b]):int(row_start[b]) + T] = xs[b].T
    word_seq = torch.full((total_words,), -1, dtype=torch.int32)
    word_lo = torch.zeros(total_words, dtype=torch.int32)
    word_hi = torch.zeros(total_words, dtype=torch.int32)
    for b in range(B):
        s = int(word_row_start[b])
        word_seq[s:s + wmax] = b
        word_lo[s:s + wmax] = torch.from_numpy(bounds[b, 0]).int()
        word_hi[s:s + wmax] = torch.from_numpy(bounds[b, 1]).int()
        word_lo[s + int(lengths[b]):s + wmax] = -1     # padded word slots
        word_hi[s + int(lengths[b]):s + wmax] = -1
    y = eng.pool(
        rows.cuda(), row_start, n_rows, word_seq.cuda(), word_lo.cuda(),
        word_hi.cuda(), method).cpu()
    out = np.zeros((B, C, wmax), dtype=np.float32)
    for b in range(B):
        s = int(word_row_start[b])
        out[b] = y[s:s + wmax].T.numpy()
        assert y[s - 1].abs().max() == 0
    return out


def test_pool_segment_assignment_bit_exact(eng):
    """Integer check of segmentation: pool one-hot frame indicators with `sum`
    so pooled[w][c] counts the frames of word w -> exact [lo, hi) recovery"""
    from emphases_b200 import engine
    utterances = []
    for seed in range(6):
        times, audio = oracle.synthetic_utterance(200 + seed)
        utterances.append((np.asarray(times), audio.shape[-1]))
    plan = engine.make_plan(utterances)
    views = eng.upload_plan(plan)
    # channel 0: 1.0, channel 1: frame index within the sequence, channel 2: idx^2
    x = torch.zeros(plan.total_rows, 80)
    for u in range(plan.n_seq):
        s, n = int(plan.row_start[u]), int(plan.n_rows[u])
        idx = torch.arange(n, dtype=torch.float32)
        x[s:s + n, 0] = 1
        x[s:s + n, 1] = idx
    y = eng.pool(
        x.cuda(), views['row_start'], views['n_rows'], views['word_seq'],
        views['word_lo'], views['word_hi'], 'sum').cpu()
    for u in range(plan.n_seq):
        expected = oracle.word_bounds(
            [tuple(t) for t in utterances[u][0].tolist()])
        s = int(plan.word_row_start[u])
        for j, (lo, hi) in enumerate(expected):
            hi = min(hi, int(plan.n_rows[u]))
            count = int(y[s + j, 0])
            total = int(y[s + j, 1])
            assert count == hi - lo
            assert total == (lo + hi - 1) * (hi - lo) // 2


def test_head(eng, golden):
    from emphases_b200 import _lib
    data = golden('c1')
    state = state_from_golden(data)
    weights = default_weights(state)
    lengths = [25, 1, 9]
    row_start, n_rows, total = make_rows(lengths)
    row_seq = eng.row_index(row_start, n_rows, len(lengths), total)
    generator = torch.Generator().manual_seed(8)
    x = torch.randn(total, 80, generator=generator)
    x[(row_seq < 0).cpu()] = 0
    logits, scores = eng.head(x.cuda(), row_seq, weights, _lib.HEAD_SIGMOID)
    layers = [(state['output_layer.weight'], state['output_layer.bias'], False)]
    from emphases_b200 import engine
    starts, _ = engine.packed_starts(lengths)
    for start, n in zip(starts, lengths):
        with torch.no_grad():
            expected = oracle._conv(
                x[start:start + n].T[None], *layers[0][:2])[0, 0]
        np.testing.assert_allclose(
            logits.cpu()[start:start + n].numpy(), expected.numpy(), atol=2e-6)
        np.testing.assert_allclose(
            scores.cpu()[start:start + n].numpy(),
            torch.sigmoid(expected).numpy(), atol=1e-6)


@pytest.mark.parametrize('batch_size', [None, 300, 100])
def test_forward_packed_golden(eng, golden, batch_size):
    """Whole path on the C1 example with the bundled checkpoint, fp32 mode:
    scores within 1e-5 of the reference, intermediates compared too"""
    from emphases_b200 import engine
    data = golden('c1')
    state = state_from_golden(data)
    weights = default_weights(state)
    times = np.asarray(data['times'])
    plan = engine.make_plan([(times, 160000)], batch_size)
    audio = torch.from_numpy(data['audio'])[0].cuda()
    result = eng.forward_packed(audio, plan, weights, keep=True)
    tag = 'full' if batch_size is None else f'bs{batch_size}'
    scores = torch.cat([
        result['scores'][s:s + n]
        for s, n in zip(plan.word_row_start, plan.n_words)]).cpu().numpy()
    expected = data[f'{tag}.scores'][0]
    error = np.abs(scores - expected).max()
    assert error < 1e-5, f'scores max-abs {error}'
    if batch_size is None:
        s, n = int(plan.row_start[0]), int(plan.n_rows[0])
        frames = result['frames'][s:s + n].cpu().numpy()
        np.testing.assert_allclose(
            frames, data['full.frame_embeddings'][0].T, atol=2e-5)
        ws, wn = int(plan.word_row_start[0]), int(plan.n_words[0])
        pooled = result['pooled'][ws:ws + wn].cpu().numpy()
        np.testing.assert_allclose(
            pooled, data['full.word_embeddings'][0].T, rtol=2e-6, atol=2e-4)


def test_forward_packed_ragged_batch(eng, golden):
    """Many ragged utterances in ONE packed launch == per-utterance oracle"""
    from emphases_b200 import engine
    data = golden('c1')
    state = state_from_golden(data)
    weights = default_weights(state)
    utterances, audios, all_times = [], [], []
    for seed in range(12):
        times, audio = oracle.synthetic_utterance(300 + seed)
        utterances.append((np.asarray(times), audio.shape[-1]))
        audios.append(audio)
        all_times.append(times)
    plan = engine.make_plan(utterances)
    packed = torch.zeros(plan.audio_samples)
    for offset, audio in zip(plan.audio_offsets, audios):
        packed[offset:offset + audio.shape[-1]] = audio[0]
    result = eng.forward_packed(packed.cuda(), plan, weights)
    scores = result['scores'].cpu()
    worst = 0.0
    for u in range(plan.n_seq):
        expected = oracle.from_alignment_and_audio(
            all_times[u], audios[u], state)[0]
        s, n = int(plan.word_row_start[u]), int(plan.n_words[u])
        worst = max(worst, (scores[s:s + n] - expected).abs().max().item())
    assert worst < 1e-5, worst


def test_conv_stack_bf16_tc(eng, golden):
    """tcgen05 bf16 conv stack vs the fp32 oracle (trained checkpoint weights):
    bf16 operands / fp32 accumulate -> relative error ~1e-2 on activations"""
    from emphases_b200 import _lib
    data = golden('c1')
    state = state_from_golden(data)
    weights = default_weights(state)
    generator = torch.Generator().manual_seed(4)
    lengths = [300, 1, 2, 131, 57, 640, 1000, 77]
    row_start, n_rows, total = make_rows(lengths)
    row_seq = eng.row_index(row_start, n_rows, len(lengths), total)
    x = torch.randn(total, 80, generator=generator)
    x[(row_seq < 0).cpu()] = 0
    layers = [(state['input_layer.weight'], state['input_layer.bias'], False)]
    layers += [
        (state[f'frame_encoder.{2 * i}.weight'],
         state[f'frame_encoder.{2 * i}.bias'], True) for i in range(6)]
    y = eng.conv_stack(x.cuda(), row_seq, weights.frame, _lib.PREC_BF16_TC).cpu()
    expected = oracle_conv_rows(state, layers, x, lengths)
    scale = expected.abs().max().item()
    error = (y - expected).abs().max().item()
    assert error < 3e-2 * scale, (error, scale)
    assert y[(row_seq < 0).cpu()].abs().max() == 0
    # bf16-emulating oracle: quantise operands like the kernel does -> tight
    emulated = torch.zeros_like(expected)
    from emphases_b200 import engine
    starts, _ = engine.packed_starts(lengths)
    for start, n in zip(starts, lengths):
        h = x[start:start + n].T[None]
        for index, (weight, bias, relu) in enumerate(layers):
            hq = h.to(torch.bfloat16).double()
            wq = weight.to(torch.bfloat16).double()
            h = (torch.nn.functional.conv1d(hq, wq, None, padding=1)
                 + bias.double()[None, :, None]).float()
            if relu:
                h = torch.relu(h)
        emulated[start:start + n] = h[0].T
    tight = (y - emulated).abs().max().item()
    # one bf16 rounding flip of an intermediate (ulp 2^-9 at 0.5) moves the output by ~3e-4
    assert tight < 2e-3 * max(scale, 1.0), (tight, scale)


def test_forward_packed_bf16_golden(eng, golden):
    """Whole path in bf16 tensor-core mode: scores within 2e-3 of the
    reference's fp32 forward (trained checkpoint, sum pooling)"""
    from emphases_b200 import _lib, engine
    data = golden('c1')
    state = state_from_golden(data)
    weights = default_weights(state)
    times = np.asarray(data['times'])
    plan = engine.make_plan([(times, 160000)], None)
    audio = torch.from_numpy(data['audio'])[0].cuda()
    result = eng.forward_packed(
        audio, plan, weights, precision=_lib.PREC_BF16_TC)
    s, n = int(plan.word_row_start[0]), int(plan.n_words[0])
    scores = result['scores'][s:s + n].cpu().numpy()
    error = np.abs(scores - data['full.scores'][0]).max()
    assert error < 2e-3, f'bf16 scores max-abs {error}'


def test_logmel_spectral_tilt_precision(eng):
    """Speech-like audio (strong low-frequency tilt): bins k and 512-k share a
    butterfly in the packed real FFT, so check the weak high-frequency bands
    against an fp64 reference -- our error must stay in the class of
    torch.stft's own fp32 error"""
    from emphases_b200 import engine
    generator = torch.Generator().manual_seed(7)
    noise = torch.randn(1, 48000, generator=generator, dtype=torch.float64)
    audio = torch.zeros_like(noise)
    state = 0.0
    values = noise[0].tolist()
    out = []
    for v in values:                       # one-pole low-pass, ~45 dB tilt
        state = 0.985 * state + v
        out.append(state)
    audio[0] = torch.tensor(out, dtype=torch.float64)
    audio = (0.5 * audio / audio.abs().max()).float()
    times = [(0.0, 1.5), (1.5, 3.0)]
    # fp64 reference of the same pipeline
    padded = torch.nn.functional.pad(audio.double(), (432, 432))
    chunk = torch.nn.functional.pad(padded, (432, 432), mode='reflect')
    stft = torch.stft(
        chunk, 1024, hop_length=160, window=torch.hann_window(1024, dtype=torch.float64),
        center=False, return_complex=True)[0]
    spectrogram = torch.sqrt(stft.real ** 2 + stft.imag ** 2 + 1e-6)
    basis = torch.from_numpy(engine.mel_basis()).double()
    exact = torch.log(torch.clamp(basis @ spectrogram, min=1e-5)).T
    reference = list(oracle.preprocess(times, audio))[0][0][0].T.double()
    plan = engine.make_plan([(np.asarray(times), 48000)])
    views = eng.upload_plan(plan)
    row_seq = eng.row_index(
        views['row_start'], views['n_rows'], plan.n_seq, plan.total_rows)
    ours = eng.logmel(audio[0].cuda(), views, plan, row_seq).cpu()[1:-1].double()
    frames = min(len(exact), len(ours))
    torch_error = (reference[:frames] - exact[:frames]).abs().max().item()
    our_error = (ours[:frames] - exact[:frames]).abs().max().item()
    assert exact.max() - exact.min() > 6          # a real dynamic range
    assert our_error < 4 * torch_error + 2e-5, (our_error, torch_error)


@pytest.mark.parametrize('mode', ['bf16x3', 'bf16x6'])
@pytest.mark.parametrize('which', ['frame', 'word'])
def test_conv_stack_bf16x3_tc(eng, golden, which, mode):
    """Split-operand modes on tcgen05: bf16x3 (hi/lo, 16 mantissa bits per
    operand) and bf16x6 (hi/mid/lo, all 24: as tight as the FFMA kernel)"""
    from emphases_b200 import _lib
    data = golden('c1')
    state = state_from_golden(data)
    weights = default_weights(state)
    generator = torch.Generator().manual_seed(5)
    lengths = [300, 1, 2, 131, 57, 640, 1000, 77]
    row_start, n_rows, total = make_rows(lengths)
    row_seq = eng.row_index(row_start, n_rows, len(lengths), total)
    x = torch.randn(total, 80, generator=generator)
    x[(row_seq < 0).cpu()] = 0
    if which == 'frame':
        layers = [(state['input_layer.weight'], state['input_layer.bias'], False)]
        layers += [
            (state[f'frame_encoder.{2 * i}.weight'],
             state[f'frame_encoder.{2 * i}.bias'], True) for i in range(6)]
        stack = weights.frame
    else:
        layers = [
            (state[f'word_decoder.{2 * i}.weight'],
             state[f'word_decoder.{2 * i}.bias'], True) for i in range(6)]
        stack = weights.word
    precision = _lib.PREC_BF16X3_TC if mode == 'bf16x3' else _lib.PREC_BF16X6_TC
    y = eng.conv_stack(x.cuda(), row_seq, stack, precision).cpu()
    expected = oracle_conv_rows(state, layers, x, lengths)
    scale = expected.abs().max().item()
    error = (y - expected).abs().max().item()
    tolerance = 5e-5 if mode == 'bf16x3' else 2e-6       # 2e-6: the fp32 kernel's own bar
    assert error < tolerance * max(scale, 1.0), (error, scale)
    assert y[(row_seq < 0).cpu()].abs().max() == 0


@pytest.mark.parametrize('batch_size', [None, 300])
def test_forward_packed_bf16x6_golden(eng, golden, batch_size):
    """Whole path with the bf16x6 conv stacks: scores within 1e-5 of the
    reference's fp32 forward, the exact mode's tolerance, on tensor cores"""
    from emphases_b200 import _lib, engine
    data = golden('c1')
    state = state_from_golden(data)
    weights = default_weights(state)
    times = np.asarray(data['times'])
    plan = engine.make_plan([(times, 160000)], batch_size)
    audio = torch.from_numpy(data['audio'])[0].cuda()
    result = eng.forward_packed(
        audio, plan, weights, precision=_lib.PREC_BF16X6_TC)
    tag = 'full' if batch_size is None else f'bs{batch_size}'
    scores = torch.cat([
        result['scores'][s:s + n]
        for s, n in zip(plan.word_row_start, plan.n_words)]).cpu().numpy()
    error = np.abs(scores - data[f'{tag}.scores'][0]).max()
    assert error < 1e-5, f'bf16x6 scores max-abs {error}'


@pytest.mark.parametrize('batch_size', [None, 300])
def test_forward_packed_bf16x3_golden(eng, golden, batch_size):
    """Whole path with the bf16x3 tensor-core conv stacks: scores within 2e-5
    (measured 5.7e-6) of the reference's fp32 forward (trained checkpoint, sum
    pooling)"""
    from emphases_b200 import _lib, engine
    data = golden('c1')
    state = state_from_golden(data)
    weights = default_weights(state)
    times = np.asarray(data['times'])
    plan = engine.make_plan([(times, 160000)], batch_size)
    audio = torch.from_numpy(data['audio'])[0].cuda()
    result = eng.forward_packed(
        audio, plan, weights, precision=_lib.PREC_BF16X3_TC)
    tag = 'full' if batch_size is None else f'bs{batch_size}'
    scores = torch.cat([
        result['scores'][s:s + n]
        for s, n in zip(plan.word_row_start, plan.n_words)]).cpu().numpy()
    error = np.abs(scores - data[f'{tag}.scores'][0]).max()
    assert error < 2e-5, f'bf16x3 scores max-abs {error}'


###############################################################################
# Word pooling fused into the tensor-core conv stack (emph_conv_stack_pool)
###############################################################################


def _fused_corpus(seeds):
    from emphases_b200 import engine
    utterances = []
    for seed in seeds:
        times, audio = oracle.synthetic_utterance(seed)
        utterances.append((np.asarray(times), audio.shape[-1]))
    return utterances, engine.make_plan(utterances)


@pytest.mark.parametrize('method', ['sum', 'average', 'max', 'center'])
@pytest.mark.parametrize('mode', ['bf16', 'bf16x3', 'bf16x6'])
def test_fused_pooling_matches_separate_kernels(eng, golden, mode, method):
    """emph_conv_stack_pool == emph_conv_stack followed by emph_pool_words:
    max / center bit for bit (the same conv rows, an exact reduction), sum /
    average within 2e-6 of the largest value (64-bit fixed-point sums against
    the pooling kernel's fp32 running sums)"""
    from emphases_b200 import _lib
    precision = {'bf16': _lib.PREC_BF16_TC, 'bf16x3': _lib.PREC_BF16X3_TC,
                 'bf16x6': _lib.PREC_BF16X6_TC}[mode]
    weights = default_weights(state_from_golden(golden('c1')))
    _, plan = _fused_corpus(range(300, 309))
    assert plan.words_disjoint()
    views = eng.upload_plan(plan)
    row_seq = eng.row_index(views['row_start'], views['n_rows'], plan.n_seq, plan.total_rows)
    generator = torch.Generator().manual_seed(3)
    x = torch.randn(plan.total_rows, 80, generator=generator).cuda()
    x[(row_seq < 0)] = 0
    frames = eng.conv_stack(x, row_seq, weights.frame, precision)
    separate = eng.pool(
        frames, views['row_start'], views['n_rows'], views['word_seq'],
        views['word_lo'], views['word_hi'], method)
    fused, kept = eng.conv_stack_pool(
        x, row_seq, weights.frame, precision, views, plan.total_word_rows, method,
        keep_frames=True)
    assert torch.equal(kept, frames)
    keep = torch.from_numpy(plan.word_seq >= 0).cuda()
    assert (fused[~keep] == 0).all()
    if method in ('max', 'center'):
        assert torch.equal(fused[keep], separate[keep])
    else:
        scale = separate[keep].abs().max().item()
        assert (fused[keep] - separate[keep]).abs().max().item() < 2e-6 * scale
    # and without the frame rows
    alone, none = eng.conv_stack_pool(
        x, row_seq, weights.frame, precision, views, plan.total_word_rows, method)
    assert none is None and torch.equal(alone, fused)


def test_fused_pooling_segment_assignment_bit_exact(eng):
    """Segmentation through the fused kernel, as integers: a one-layer identity
    stack in the exact bf16x6 mode passes frame indicators / frame indices
    through unchanged, so pooled sums recover every word's [lo, hi) exactly"""
    from emphases_b200 import _lib, engine
    utterances, plan = _fused_corpus(range(200, 206))
    views = eng.upload_plan(plan)
    row_seq = eng.row_index(views['row_start'], views['n_rows'], plan.n_seq, plan.total_rows)
    weights = torch.zeros(1, 3, 80, 80)
    weights[0, 1] = torch.eye(80)
    stack = engine.ConvStack(
        weights.cuda(), torch.zeros(1, 80).cuda(),
        np.asarray([_lib.ACT_NONE], dtype=np.int32), 3, 80)
    x = torch.zeros(plan.total_rows, 80)
    for u in range(plan.n_seq):
        s, n = int(plan.row_start[u]), int(plan.n_rows[u])
        x[s:s + n, 0] = 1
        x[s:s + n, 1] = torch.arange(n, dtype=torch.float32)
    pooled, _ = eng.conv_stack_pool(
        x.cuda(), row_seq, stack, _lib.PREC_BF16X6_TC, views, plan.total_word_rows, 'sum')
    pooled = pooled.cpu()
    center, _ = eng.conv_stack_pool(
        x.cuda(), row_seq, stack, _lib.PREC_BF16X6_TC, views, plan.total_word_rows, 'center')
    center = center.cpu()
    for u in range(plan.n_seq):
        expected = oracle.word_bounds([tuple(t) for t in utterances[u][0].tolist()])
        s = int(plan.word_row_start[u])
        for j, (lo, hi) in enumerate(expected):
            hi = min(hi, int(plan.n_rows[u]))
            assert pooled[s + j, 0].item() == hi - lo
            assert pooled[s + j, 1].item() == (lo + hi - 1) * (hi - lo) // 2
            assert center[s + j, 1].item() == (lo + hi) // 2


def test_fused_pooling_is_deterministic_and_packing_stable(eng, golden):
    """The integer atomics make the fused sums bit-identical run to run whatever
    order the partial sums land in; packing the same utterances behind a
    different prefix moves the cuts between the <= 16-row fp32 partial sums, so
    those results agree to fp32 rounding of the sums (not bit for bit)"""
    from emphases_b200 import _lib
    weights = default_weights(state_from_golden(golden('c1')))
    results = []
    for seeds in ([400, 401, 402, 403], [400, 401, 402, 403], [410, 400, 401, 402, 403]):
        _, plan = _fused_corpus(seeds)
        views = eng.upload_plan(plan)
        row_seq = eng.row_index(
            views['row_start'], views['n_rows'], plan.n_seq, plan.total_rows)
        x = torch.zeros(plan.total_rows, 80)
        for u, seed in enumerate(seeds):
            s, n = int(plan.row_start[u]), int(plan.n_rows[u])
            generator = torch.Generator().manual_seed(seed)
            x[s:s + n] = torch.randn(n, 80, generator=generator)
        pooled, _ = eng.conv_stack_pool(
            x.cuda(), row_seq, weights.frame, _lib.PREC_BF16X6_TC, views,
            plan.total_word_rows, 'sum')
        first = len(seeds) - 4
        start = int(plan.word_row_start[first])
        results.append(pooled[start:].cpu())
    assert torch.equal(results[0], results[1])
    scale = results[0].abs().max().item()
    assert (results[0] - results[2]).abs().max().item() < 1e-6 * max(scale, 1.0)


###############################################################################
# Sample-rate conversion fused into the log-mel front end
###############################################################################


@pytest.mark.parametrize('dtype', ['f32', 'i16'])
@pytest.mark.parametrize('rate', [44100, 24000, 8000, 22050])
def test_logmel_with_fused_resampling_equals_resampling_first(eng, rate, dtype):
    """emph_logmel_resampled_* on source-rate audio == emph_resample_* followed
    by emph_logmel_* (bit for bit: same filter bank, same summation order), on
    a ragged packed corpus with chunk edges, for fp32 and int16 PCM sources"""
    from emphases_b200 import engine, resampling, scheduler
    device = torch.device('cuda', 0)
    generator = torch.Generator().manual_seed(rate)
    sources, audios, utterances = [], [], []
    for index, seconds in enumerate([0.37, 1.93, 0.8, 2.71, 0.05]):
        samples = int(seconds * rate) + index
        source = (0.1 * torch.randn(1, samples, generator=generator)).clamp(-1, 1)
        if dtype == 'i16':
            source = (source * 32768.).round().clamp(-32768, 32767).to(torch.int16)
        as_float = source.float() / 32768. if dtype == 'i16' else source
        audio = resampling.resample(as_float, rate, 16000, device).cpu()
        duration = audio.shape[-1] / 16000.
        words = max(1, int(3 * duration))
        edges = np.linspace(0., duration, words + 1)
        utterances.append((np.stack([edges[:-1], edges[1:]], 1), audio.shape[-1]))
        sources.append(source)
        audios.append(audio)
    plan = engine.make_plan(utterances)
    views = eng.upload_plan(plan)
    row_seq = eng.row_index(views['row_start'], views['n_rows'], plan.n_seq, plan.total_rows)
    packed16 = torch.zeros(plan.audio_samples)
    for offset, audio in zip(plan.audio_offsets, audios):
        packed16[offset:offset + audio.shape[-1]] = audio[0]
    expected = eng.logmel(packed16.cuda(), views, plan, row_seq).clone()

    packed_source = scheduler.pack_audio(sources, dtype=sources[0].dtype, pin=False)
    lengths16 = resampling.resampled_lengths(packed_source.lengths, rate, 16000)
    assert [int(n) for n in lengths16] == [a.shape[-1] for a in audios]
    resample = {
        'bank': resampling.device_bank(rate, 16000, device),
        'source_off': torch.from_numpy(
            packed_source.offsets[plan.utterance].astype(np.int64)).cuda(),
        'source_len': torch.from_numpy(
            packed_source.lengths[plan.utterance].astype(np.int32)).cuda()}
    fused = eng.logmel(
        packed_source.buffer.cuda(), views, plan, row_seq, resample=resample)
    assert torch.equal(fused, expected)


def _attention_case(seed, lengths, keys, channels=80, spread=1.0):
    """Packed q, k, v rows of ragged sequences (with padded, key-masked tails)
    and the fp64 attention of every head computed with torch"""
    from emphases_b200 import transformer
    heads = transformer.HEADS
    rng = np.random.default_rng(seed)
    row_start, n_rows, total = make_rows(lengths)
    q, k, v = (
        torch.from_numpy(rng.standard_normal((total, channels)).astype(np.float32) * scale)
        for scale in (spread, spread, 1.0))
    head_dim = channels // heads
    expected = torch.zeros(total, channels, dtype=torch.float64)
    for start, length, valid in zip(row_start.tolist(), lengths, keys):
        for head in range(heads):
            columns = slice(head * head_dim, (head + 1) * head_dim)
            logits = q[start:start + length, columns].double() @ \
                k[start:start + valid, columns].double().T / np.sqrt(head_dim)
            expected[start:start + length, columns] = \
                torch.softmax(logits, dim=1) @ v[start:start + valid, columns].double()
    return row_start, n_rows, total, q, k, v, expected


def _run_attention(eng, row_start, n_rows, total, q, k, v, lengths, keys, mode):
    from emphases_b200 import _lib, transformer
    channels = q.shape[1]
    block_seq, block_q0 = transformer.query_blocks(np.asarray(lengths))
    device = torch.device('cuda:0')
    row_seq = eng.row_index(row_start, n_rows, len(lengths), total)
    n_keys = torch.tensor(keys, dtype=torch.int32, device=device)
    d_seq = torch.from_numpy(block_seq).to(device)
    d_q0 = torch.from_numpy(block_q0).to(device)
    q, k, v = q.to(device), k.to(device), v.to(device)
    out = torch.full((total, channels), float('nan'), device=device)
    scale = 1.0 / np.sqrt(channels // transformer.HEADS)
    if mode is None:
        _lib.call(
            'emph_attention_rows', _lib.ptr(q), _lib.ptr(k), _lib.ptr(v), channels,
            transformer.HEADS, _lib.ptr(row_start), _lib.ptr(n_rows), _lib.ptr(n_keys),
            _lib.ptr(row_seq), total, _lib.ptr(d_seq), _lib.ptr(d_q0), len(block_seq),
            scale, _lib.ptr(out), _lib.stream_ptr())
    else:
        workspace = transformer.attention_workspace(total, channels, mode, device)
        workspace.fill_(0xff)                  # NaN patterns wherever staging skips a byte
        _lib.call(
            'emph_attention_rows_tc', _lib.ptr(q), _lib.ptr(k), _lib.ptr(v), channels,
            transformer.HEADS, _lib.ptr(row_start), _lib.ptr(n_rows), _lib.ptr(n_keys),
            _lib.ptr(row_seq), total, _lib.ptr(d_seq), _lib.ptr(d_q0), len(block_seq),
            scale, mode, _lib.ptr(workspace), workspace.numel(), _lib.ptr(out),
            _lib.stream_ptr())
    torch.cuda.synchronize()
    return out.cpu()


# max-abs error of the attention output (unit-variance q, k, v: logits far
# larger than the model's) vs fp64, per form: fp32 CUDA cores, split bf16, fp16
ATTENTION_TOLERANCE = {None: 2e-6, 1: 3e-5, 0: 4e-3}


@pytest.mark.parametrize('channels', [80, 64, 128])
@pytest.mark.parametrize('mode', [None, 1, 0], ids=['fp32', 'bf16x3', 'fp16'])
def test_attention_rows_all_forms(eng, mode, channels):
    """Block-diagonal, key-masked attention over ragged packed sequences: the
    CUDA-core kernel and the tensor-core forms against fp64 torch; sequence
    lengths around the 64-key tile and the 128-query block, single rows, padded
    (masked) tails, head dims 40 / 32 / 64"""
    lengths = [1, 7, 63, 64, 65, 127, 128, 129, 200, 333, 700]
    keys = [1, 3, 63, 64, 60, 127, 128, 1, 137, 333, 641]
    row_start, n_rows, total, q, k, v, expected = _attention_case(
        5, lengths, keys, channels)
    got = _run_attention(eng, row_start, n_rows, total, q, k, v, lengths, keys, mode)
    rows = torch.cat([
        torch.arange(start, start + length)
        for start, length in zip(row_start.tolist(), lengths)])
    separators = torch.ones(total, dtype=torch.bool)
    separators[rows] = False
    assert torch.all(got[separators] == 0)
    error = (got[rows].double() - expected[rows]).abs().max().item()
    print(f'attention mode {mode}, {channels} channels: max error {error:.3e}')
    assert error < ATTENTION_TOLERANCE[mode], error


def test_attention_tensor_core_peaked_softmax(eng):
    """Large logits (|q.k|/sqrt(d) up to ~40: one key dominates each row): the
    split form still follows fp64 to 3e-4 where 16-bit logits would not"""
    lengths, keys = [300, 90], [300, 77]
    row_start, n_rows, total, q, k, v, expected = _attention_case(
        9, lengths, keys, spread=2.5)
    rows = torch.cat([
        torch.arange(start, start + length)
        for start, length in zip(row_start.tolist(), lengths)])
    exact = _run_attention(eng, row_start, n_rows, total, q, k, v, lengths, keys, None)
    split = _run_attention(eng, row_start, n_rows, total, q, k, v, lengths, keys, 1)
    assert (exact[rows].double() - expected[rows]).abs().max() < 2e-5
    assert (split[rows].double() - expected[rows]).abs().max() < 3e-4


@pytest.mark.parametrize('parts', [1, 2, 3])
def test_transformer_fused_passes(eng, parts):
    """csrc/transformer_tc.cu against fp64 torch on ragged packed rows: the
    out-projection + residual + LayerNorm pass, the feed-forward + residual +
    LayerNorm pass, and the q / k / v pass whose k / v records feed
    emph_attention_rows_staged (checked through the attention output)"""
    from emphases_b200 import _lib, transformer
    device = torch.device('cuda:0')
    lengths, keys = [1, 37, 128, 130, 261], [1, 30, 128, 97, 261]
    row_start, n_rows, total = make_rows(lengths)
    row_seq = eng.row_index(row_start, n_rows, len(lengths), total)
    rows = torch.cat([
        torch.arange(start, start + length)
        for start, length in zip(row_start.tolist(), lengths)])
    generator = torch.Generator().manual_seed(7)
    channels = 80

    def normal(*shape, scale=1.0):
        return torch.randn(*shape, generator=generator) * scale

    x, residual = normal(total, channels), normal(total, channels)
    weight = [normal(channels, channels, scale=channels ** -.5) for _ in range(5)]
    bias = [normal(channels, scale=.1) for _ in range(5)]
    gamma, beta = 1 + normal(channels, scale=.1), normal(channels, scale=.1)
    tolerance = {1: 4e-3, 2: 3e-5, 3: 2e-6}[parts]      # one fp16 value / 2 / 3 bf16 parts

    def layernorm(value):
        return torch.nn.functional.layer_norm(
            value, (channels,), gamma.double(), beta.double(), 1e-5)

    def check(got, want):
        got = got.cpu()
        separators = torch.ones(total, dtype=torch.bool)
        separators[rows] = False
        assert torch.all(got[separators] == 0)
        error = (got[rows].double() - want[rows]).abs().max().item()
        assert error < tolerance, error

    dx, dres = x.to(device), residual.to(device)
    dgamma, dbeta = gamma.to(device), beta.to(device)
    stream = _lib.stream_ptr()

    # y = LayerNorm(residual + x W^T + b)
    blob = transformer.split_parts([weight[0]], parts, device)
    dbias = bias[0].to(device)
    y = torch.full_like(dx, float('nan'))
    _lib.call(
        'emph_transformer_proj_norm', _lib.ptr(dx), _lib.ptr(dres), total, channels,
        _lib.ptr(blob), _lib.ptr(dbias), parts, _lib.ptr(dgamma), _lib.ptr(dbeta), 1e-5,
        _lib.ptr(row_seq), _lib.ptr(y), stream)
    check(y, layernorm(residual.double() + x.double() @ weight[0].double().T + bias[0].double()))

    # y = LayerNorm(x + relu(x W1^T + b1) W2^T + b2)
    blob = transformer.split_parts(weight[:2], parts, device)
    dbias = torch.cat(bias[:2]).to(device)
    y = torch.full_like(dx, float('nan'))
    _lib.call(
        'emph_transformer_ffn_norm', _lib.ptr(dx), total, channels, _lib.ptr(blob),
        _lib.ptr(dbias), parts, _lib.ptr(dgamma), _lib.ptr(dbeta), 1e-5, _lib.ptr(row_seq),
        _lib.ptr(y), stream)
    hidden = torch.relu(x.double() @ weight[0].double().T + bias[0].double())
    check(y, layernorm(x.double() + hidden @ weight[1].double().T + bias[1].double()))

    # the two passes above as one: LayerNorm(n + relu(n W1^T + b1) W2^T + b2) with
    # n = LayerNorm(residual + x W0^T + b0) (second LayerNorm with swapped vectors)
    blob = transformer.split_parts(weight[:3], parts, device)
    dbias = torch.cat(bias[:3]).to(device)
    y = torch.full_like(dx, float('nan'))
    _lib.call(
        'emph_transformer_layer_tail', _lib.ptr(dx), _lib.ptr(dres), total, channels,
        _lib.ptr(blob), _lib.ptr(dbias), parts, _lib.ptr(dgamma), _lib.ptr(dbeta),
        _lib.ptr(dbeta), _lib.ptr(dgamma), 1e-5, _lib.ptr(row_seq), _lib.ptr(y), stream)
    first = layernorm(residual.double() + x.double() @ weight[0].double().T + bias[0].double())
    hidden = torch.relu(first @ weight[1].double().T + bias[1].double())
    second = torch.nn.functional.layer_norm(
        first + hidden @ weight[2].double().T + bias[2].double(), (channels,),
        beta.double(), gamma.double(), 1e-5)
    check(y, second)

    # q rows + k / v records -> attention
    q64, k64, v64 = (
        x.double() @ weight[2 + i].double().T + bias[2 + i].double() for i in range(3))
    blob = transformer.split_parts(weight[2:], parts, device)
    dbias = torch.cat(bias[2:]).to(device)
    heads, head_dim = transformer.HEADS, channels // transformer.HEADS
    expected = torch.zeros(total, channels, dtype=torch.float64)
    for start, length, valid in zip(row_start.tolist(), lengths, keys):
        for head in range(heads):
            columns = slice(head * head_dim, (head + 1) * head_dim)
            logits = q64[start:start + length, columns] @ \
                k64[start:start + valid, columns].T / np.sqrt(head_dim)
            expected[start:start + length, columns] = \
                torch.softmax(logits, dim=1) @ v64[start:start + valid, columns]
    block_seq, block_q0 = transformer.query_blocks(np.asarray(lengths))
    d_seq, d_q0 = torch.from_numpy(block_seq).to(device), torch.from_numpy(block_q0).to(device)
    n_keys = torch.tensor(keys, dtype=torch.int32, device=device)
    for mode, bound in ((1, 2e-4 if parts > 1 else 2e-2), (0, 2e-2)):
        records = transformer.attention_workspace(total, channels, mode, device, zero=True)
        q = torch.full_like(dx, float('nan'))
        out = torch.zeros_like(dx)
        _lib.call(
            'emph_transformer_qkv', _lib.ptr(dx), total, channels, _lib.ptr(blob),
            _lib.ptr(dbias), parts, mode, _lib.ptr(q), _lib.ptr(records), records.numel(),
            stream)
        _lib.call(
            'emph_attention_rows_staged', _lib.ptr(q), _lib.ptr(records), records.numel(),
            channels, heads, _lib.ptr(row_start), _lib.ptr(n_rows), _lib.ptr(n_keys), total,
            _lib.ptr(d_seq), _lib.ptr(d_q0), len(block_seq), 1.0 / np.sqrt(head_dim), mode,
            _lib.ptr(out), stream)
        torch.cuda.synchronize()
        assert (q.cpu()[rows].double() - q64[rows]).abs().max() < tolerance
        error = (out.cpu()[rows].double() - expected[rows]).abs().max().item()
        assert error < bound, (mode, error)
