"""CPU oracle for the emphases batched-inference hot path.

ORACLE / TEST INFRASTRUCTURE.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` legs may import this module.
The product package `emphases_b200/` must never import it.

This is a restatement, in plain torch-CPU fp32 ops and numpy, of the algorithm
in interactiveaudiolab/emphases for the path
    waveform + word alignment -> log-mel -> Conv1d/ReLU (or Transformer) frame
    encoder -> word-bound segment pooling -> word decoder -> Conv1d(->1) ->
    sigmoid.
Each function cites the reference file:line it follows (paths relative to the
reference root).  It uses the same library primitives the reference calls
(`torch.stft`, `F.conv1d`, `F.multi_head_attention_forward`-equivalent
`nn.TransformerEncoderLayer` math), so it tracks the reference to float
rounding.

PINNING: the reference ships no tests or golden vectors.
This oracle is pinned against outputs of the UNMODIFIED reference run in the
build container through `oracle/ref_stubs.py`; `oracle/gen_golden.py` writes
those outputs to `tests/golden/*.npz` and `tests/test_oracle_golden.py` checks
this file against them.  Two third-party pieces remain "parity unpinned"
because their source is not in the container:
  * `librosa.filters.mel` (reference call site emphases/data/preprocess/
    mels.py:97-100; unpinned version, setup.py:15-31) -- `mel_basis` restates
    the published Slaney-scale / Slaney-norm algorithm.  The basis is an
    explicit INPUT to both this oracle and the CUDA kernels, so kernel parity
    does not depend on it.
  * `pypar.Alignment` slicing / `word_bounds` (call sites emphases/core.py:
    384-390) -- restated as "re-base slice to its first word's start, then
    int(t * sr / hop)".
  * `torchutil.metrics.{Average, MeanStd, PearsonCorrelation}` (call sites
    emphases/evaluate/metrics.py:15-18,59,82,103) -- restated in `evaluate`
    (running average; Welford mean / SAMPLE std over python floats; Pearson =
    sum((p - mean_p)(t - mean_t)) / (n std_p std_t)).  The reference's own
    metric subclasses run unmodified on the same restatement in
    `oracle/ref_stubs.py` to produce tests/golden/evaluate.npz.
"""
import math

import numpy as np
import torch

###############################################################################
# Constants (emphases/config/defaults.py:53-74)
###############################################################################

SAMPLE_RATE = 16000
HOPSIZE = 160
NUM_FFT = 1024
WINDOW_SIZE = 1024
NUM_MELS = 80


def default_config():
    """Model/config constants read by the path (emphases/config/defaults.py)"""
    return dict(
        ARCHITECTURE='convolution',      # defaults.py:184
        CHANNELS=80,                     # :187
        DECODER_KERNEL_SIZE=3,           # :190
        DROPOUT=None,                    # :193
        DOWNSAMPLE_LOCATION='intermediate',  # :197
        DOWNSAMPLE_METHOD='sum',         # :201
        ENCODER_KERNEL_SIZE=3,           # :204
        LAYERS=6,                        # :207
        ACTIVATION='relu',               # :181 (torch.nn.ReLU)
        LOSS='bce',                      # :227
        NORMALIZE=False,                 # :104
        NUM_FEATURES=80)                 # static.py:42-46


###############################################################################
# librosa.filters.mel restatement (third party; call site mels.py:97-100)
###############################################################################


def _hz_to_mel(frequencies):
    frequencies = np.asanyarray(frequencies, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = frequencies / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    if frequencies.ndim:
        log_t = frequencies >= min_log_hz
        mels[log_t] = min_log_mel + np.log(frequencies[log_t] / min_log_hz) / logstep
    elif frequencies >= min_log_hz:
        mels = min_log_mel + np.log(frequencies / min_log_hz) / logstep
    return mels


def _mel_to_hz(mels):
    mels = np.asanyarray(mels, dtype=np.float64)
    f_sp = 200.0 / 3
    freqs = f_sp * mels
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    log_t = mels >= min_log_mel
    freqs[log_t] = min_log_hz * np.exp(logstep * (mels[log_t] - min_log_mel))
    return freqs


def mel_basis(sr=SAMPLE_RATE, n_fft=NUM_FFT, n_mels=NUM_MELS):
    """Slaney mel filterbank, float32 (n_mels, 1 + n_fft // 2)"""
    fmin, fmax = 0.0, sr / 2.0
    weights = np.zeros((n_mels, 1 + n_fft // 2), dtype=np.float32)
    fftfreqs = np.fft.rfftfreq(n=n_fft, d=1.0 / sr)
    mel_f = _mel_to_hz(
        np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]
    return weights


###############################################################################
# Time conversion (emphases/convert.py:9-36)
###############################################################################


def seconds_to_frames(seconds):
    """convert.py:19-31: (seconds * SAMPLE_RATE) // HOPSIZE, float floor-div"""
    return (seconds * SAMPLE_RATE) // HOPSIZE


def word_bounds(times):
    """pypar.Alignment[start:end].word_bounds(16000, 160, silences=True)

    `times`: list of (start, end) seconds of the words of one chunk, in
    ABSOLUTE time; the slice is re-based to its first word's start.
    """
    origin = times[0][0]
    return [
        (int((start - origin) * SAMPLE_RATE / HOPSIZE),
         int((end - origin) * SAMPLE_RATE / HOPSIZE))
        for start, end in times]


###############################################################################
# Feature extraction (emphases/data/preprocess/mels.py:16-59,94-109)
###############################################################################


def logmel(audio, basis=None, normalize=False):
    """mels.from_audio: audio (1, L) float32 -> (80, L // 160) float32

    Raises RuntimeError (from the reflect pad) when L <= 432, as the
    reference does (core.py:413-415 swallows it).
    """
    window = torch.hann_window(WINDOW_SIZE, dtype=audio.dtype)   # mels.py:24
    size = (NUM_FFT - HOPSIZE) // 2                              # :32
    audio = torch.nn.functional.pad(audio, (size, size), mode='reflect')
    stft = torch.stft(                                           # :39-47
        audio.squeeze(1),
        NUM_FFT,
        hop_length=HOPSIZE,
        window=window,
        center=False,
        normalized=False,
        onesided=True,
        return_complex=True)
    stft = torch.view_as_real(stft)[0]                           # :48
    spectrogram = torch.sqrt(stft.pow(2).sum(-1) + 1e-6)         # :51
    if basis is None:
        basis = mel_basis()
    basis = torch.as_tensor(basis).to(spectrogram.dtype)
    mels = torch.log(torch.clamp(torch.matmul(basis, spectrogram), min=1e-5))
    if normalize:                                                # :57-58
        return (mels + 10.) / 10.
    return mels


###############################################################################
# Chunker (emphases/core.py:345-418)
###############################################################################


def chunk_plan(times, num_samples, batch_size=None):
    """Integer plan of emphases.preprocess: one dict per yielded chunk.

    times: list of (start, end) seconds for every word of the utterance.
    num_samples: audio length at 16 kHz (before the 432-sample zero pad).
    Returns list of dict(word_start, word_end, start_sample, end_sample,
    length, frames, bounds) in PADDED-audio sample coordinates; chunks whose
    length is <= 432 samples are dropped (core.py:413-415).
    """
    padding = int((WINDOW_SIZE - HOPSIZE) / 2)                   # core.py:357
    padded = num_samples + 2 * padding
    total_frames = int(padded / HOPSIZE)                         # :359
    batch_size = total_frames if batch_size is None else batch_size
    plan = []
    start = 0
    while start < len(times):                                    # :365
        frames = 0.
        end = start + 1
        while end < len(times):                                  # :370
            duration = times[end - 1][1] - times[end - 1][0]
            frames += seconds_to_frames(duration)                # :373-374
            if int(frames) > batch_size:                         # :377
                break
            end += 1
        bounds = word_bounds(times[start:end])                   # :384-392
        start_sample = int(HOPSIZE * int(seconds_to_frames(times[start][0])))
        end_sample = int(HOPSIZE * int(seconds_to_frames(times[end - 1][1])))
        lo = min(max(start_sample, 0), padded)                   # torch slice
        hi = min(max(end_sample, 0), padded)
        length = max(hi - lo, 0)
        if length > padding:                                     # :413-415
            plan.append(dict(
                word_start=start,
                word_end=end,
                start_sample=lo,
                end_sample=hi,
                length=length,
                frames=length // HOPSIZE,
                bounds=bounds))
        start = end
    return plan


def preprocess(times, audio, batch_size=None, basis=None):
    """emphases.preprocess (core.py:345-418): yields (features, word_bounds)

    audio: (C, T) float32 at 16 kHz.  features (1, 80, F) float32,
    word_bounds (1, 2, W) int64.
    """
    padding = int((WINDOW_SIZE - HOPSIZE) / 2)
    padded = torch.nn.functional.pad(audio, (padding, padding))
    for chunk in chunk_plan(times, audio.shape[-1], batch_size):
        batch_audio = padded[:, chunk['start_sample']:chunk['end_sample']]
        features = logmel(batch_audio, basis)[None]              # pre/core.py:125
        bounds = torch.tensor(chunk['bounds'], dtype=torch.long).T[None]
        yield features, bounds


###############################################################################
# Frame <-> word resampling (emphases/core.py:426-469, 552-586)
###############################################################################


def downsample(xs, word_bounds_, word_lengths, method):
    """emphases.downsample, core.py:426-469 (python double loop restated)"""
    if method in ('average', 'max', 'sum'):
        result = torch.zeros(
            (xs.shape[0], xs.shape[1], int(word_lengths.max())),
            dtype=xs.dtype)
        for i in range(xs.shape[0]):
            for j in range(int(word_lengths[i])):
                start = int(word_bounds_[i, 0, j])
                end = int(word_bounds_[i, 1, j])
                segment_ = xs[i][:, start:end]
                if method == 'average':
                    result[i, :, j] = segment_.mean(dim=1)
                elif method == 'max':
                    result[i, :, j] = segment_.max(dim=1).values
                else:
                    result[i, :, j] = segment_.sum(dim=1)
        return result
    if method == 'center':                                       # :458-466
        indices = (word_bounds_[:, 0] + word_bounds_[:, 1]) // 2
        return xs.transpose(1, 2)[
            torch.arange(xs.shape[0])[:, None], indices].transpose(1, 2)
    raise ValueError(f'Interpolation method {method} is not defined')


def segment(xs, word_bounds_, word_lengths):
    """emphases.segment, core.py:552-586"""
    max_length = int((word_bounds_[:, 1] - word_bounds_[:, 0]).max())
    batch = word_bounds_.shape[0] * word_bounds_.shape[2]
    result = torch.zeros((batch, xs.shape[1], max_length), dtype=xs.dtype)
    result_bounds = torch.zeros((batch, 2, 1), dtype=torch.long)
    result_lengths = torch.zeros((batch,), dtype=torch.long)
    for i in range(xs.shape[0]):
        words = int(word_lengths[i])
        for j in range(word_bounds_.shape[2]):
            k = min(j, words - 1)                                # :574
            start = int(word_bounds_[i, 0, k])
            end = int(word_bounds_[i, 1, k])
            frames = end - start
            index = i * word_bounds_.shape[2] + j
            result[index, :, :frames] = xs[i][:, start:end]
            result_bounds[index, 1, 0] = frames
            result_lengths[index] = frames
    return result, result_bounds, result_lengths


def upsample(xs, word_bounds_, word_lengths, frame_lengths, method='linear'):
    """emphases.upsample, core.py:472-544 (word -> frame resolution)"""
    result = torch.zeros(
        (xs.shape[0], xs.shape[1], int(frame_lengths.max())), dtype=xs.dtype)
    for i in range(xs.shape[0]):
        word_length, frame_length = int(word_lengths[i]), int(frame_lengths[i])
        x = xs[i][..., :word_length]
        word_bound = word_bounds_[i][..., :word_length]
        word_times = (
            word_bound[0] + (word_bound[1] - word_bound[0]) / 2.)[None]
        frame_times = .5 + torch.arange(frame_length)[None]
        if x.shape[1] == 1:                                      # :497-498
            result[i, :, :frame_length] = x[0]
        elif method == 'linear':                                 # :501-523
            slope = (
                (x[:, 1:] - x[:, :-1]) /
                (word_times[:, 1:] - word_times[:, :-1]))
            intercept = x[:, :-1] - slope.mul(word_times[:, :-1])
            indices = torch.sum(
                torch.ge(frame_times[:, :, None], word_times[:, None, :]),
                -1) - 1
            indices = torch.clamp(indices, 0, slope.shape[-1] - 1)
            # line_idx is all zeros in the reference (linspace(0, 1, 1)): the
            # slope/intercept of channel 0 are used for every channel
            line_idx = torch.zeros_like(indices)
            result[i, :, :frame_length] = (
                slope[line_idx, indices].mul(frame_times) +
                intercept[line_idx, indices])
        elif method == 'nearest':                                # :526-536
            indices = torch.sum(
                torch.ge(frame_times[:, :, None], word_times[:, None, :]),
                -1) - 1
            indices = torch.clamp(indices, 0, word_times.shape[-1] - 1)
            result[i, :, :frame_length] = torch.index_select(x, 1, indices[0])
        else:
            raise ValueError(
                f'Interpolation method {method} is not defined')
    return result


def mask_from_lengths(lengths):
    """model/core.py:146-149"""
    x = torch.arange(int(lengths.max()), dtype=lengths.dtype)
    return (x.unsqueeze(0) < lengths.unsqueeze(1)).unsqueeze(1)


###############################################################################
# Model (emphases/model/core.py:13-138, layers/convolution.py:13-37,
#        layers/transformer.py:13-52)
###############################################################################


_ACTIVATIONS = {
    'relu': torch.nn.functional.relu,
    'gelu': torch.nn.functional.gelu,
    'leaky_relu': torch.nn.functional.leaky_relu,
    'silu': torch.nn.functional.silu,
}


def _conv(x, weight, bias):
    """nn.Conv1d(padding='same') for odd kernels: zero pad (k-1)/2 each side"""
    return torch.nn.functional.conv1d(
        x, weight, bias, padding=(weight.shape[-1] - 1) // 2)


def _sequential_step(config):
    """Index stride in the Sequential: conv, activation(, dropout)"""
    return 3 if config.get('DROPOUT') is not None else 2     # conv.py:25-30


def convolution_stack(state, prefix, x, config):
    """layers/convolution.py:13-37: LAYERS x [Conv1d 'same' -> activation]"""
    activation = _ACTIVATIONS[config['ACTIVATION']]
    step = _sequential_step(config)
    for layer in range(config['LAYERS']):
        x = activation(_conv(
            x,
            state[f'{prefix}.{layer * step}.weight'],
            state[f'{prefix}.{layer * step}.bias']))
    return x


def positional_encoding(channels, max_len=5000):
    """layers/transformer.py:40-48"""
    index = torch.arange(max_len).unsqueeze(1)
    frequency = torch.exp(
        torch.arange(0, channels, 2) * (-math.log(10000.0) / channels))
    encoding = torch.zeros(max_len, 1, channels)
    encoding[:, 0, 0::2] = torch.sin(index * frequency)
    encoding[:, 0, 1::2] = torch.cos(index * frequency)
    return encoding


def transformer_stack(state, prefix, x, lengths, config):
    """layers/transformer.py:25-30 in eval mode (dropout = identity).

    nn.TransformerEncoderLayer(d, nhead=2, dim_feedforward=d), post-norm,
    ReLU, eps 1e-5, key-padding mask = ~mask_from_lengths(lengths).
    x: (B, C, T) -> (B, C, T)
    """
    channels = config['CHANNELS']
    heads = 2
    head_dim = channels // heads
    mask = mask_from_lengths(lengths).squeeze(1)             # (B, T) valid
    h = x.permute(2, 0, 1)                                   # (T, B, C)
    h = h + state[f'{prefix}.position.encoding'][:h.size(0)]
    T, B, _ = h.shape
    for layer in range(config['LAYERS']):
        p = f'{prefix}.model.layers.{layer}'
        qkv = torch.nn.functional.linear(
            h,
            state[f'{p}.self_attn.in_proj_weight'],
            state[f'{p}.self_attn.in_proj_bias'])
        q, k, v = qkv.chunk(3, dim=-1)

        def split(t):
            return t.reshape(T, B, heads, head_dim).permute(1, 2, 0, 3)
        q, k, v = split(q), split(k), split(v)               # (B, H, T, D)
        scores = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(head_dim)
        scores = scores.masked_fill(~mask[:, None, None, :], float('-inf'))
        attention = torch.softmax(scores, dim=-1)
        context = torch.matmul(attention, v)                 # (B, H, T, D)
        context = context.permute(2, 0, 1, 3).reshape(T, B, channels)
        context = torch.nn.functional.linear(
            context,
            state[f'{p}.self_attn.out_proj.weight'],
            state[f'{p}.self_attn.out_proj.bias'])
        h = torch.nn.functional.layer_norm(
            h + context, (channels,),
            state[f'{p}.norm1.weight'], state[f'{p}.norm1.bias'], 1e-5)
        feedforward = torch.nn.functional.linear(
            torch.relu(torch.nn.functional.linear(
                h, state[f'{p}.linear1.weight'], state[f'{p}.linear1.bias'])),
            state[f'{p}.linear2.weight'], state[f'{p}.linear2.bias'])
        h = torch.nn.functional.layer_norm(
            h + feedforward, (channels,),
            state[f'{p}.norm2.weight'], state[f'{p}.norm2.bias'], 1e-5)
    return h.permute(1, 2, 0)


def layers(state, prefix, x, lengths, config):
    """layers/__init__.py:7-14"""
    if config['ARCHITECTURE'] == 'convolution':
        return convolution_stack(state, prefix, x, config)
    if config['ARCHITECTURE'] == 'transformer':
        return transformer_stack(state, prefix, x, lengths, config)
    raise ValueError(
        f'Network layer {config["ARCHITECTURE"]} is not defined')


def model_forward(
    state,
    features,
    frame_lengths,
    word_bounds_,
    word_lengths,
    config=None,
    training=False,
    return_intermediates=False):
    """emphases.Model.forward (model/core.py:39-138), fp32, eval mode.

    state: dict of fp32 tensors keyed like the reference state_dict.
    features (B, 80, T); frame_lengths (B,); word_bounds_ (B, 2, Wmax) int64;
    word_lengths (B,).  Returns logits (B, 1, Wmax) (or (B, 1, T) for the
    'inference' location in training mode).
    """
    config = dict(default_config(), **(config or {}))
    location = config['DOWNSAMPLE_LOCATION']
    method = config['DOWNSAMPLE_METHOD']
    inter = {}

    def output_layer(x):
        return _conv(
            x, state['output_layer.weight'], state['output_layer.bias'])

    def input_layer(x):
        return _conv(
            x, state['input_layer.weight'], state['input_layer.bias'])

    if location == 'input':                                      # :41-87
        segments, bounds, lengths = segment(
            features, word_bounds_, word_lengths)
        frame_embeddings = layers(
            state, 'frame_encoder', input_layer(segments), lengths, config)
        if method == 'average':
            word_embeddings = frame_embeddings.mean(dim=2, keepdim=True)
        elif method == 'max':
            word_embeddings = frame_embeddings.max(
                dim=2, keepdim=True).values
        elif method == 'sum':
            word_embeddings = frame_embeddings.sum(dim=2, keepdim=True)
        elif method == 'center':
            word_embeddings = downsample(
                frame_embeddings,
                bounds,
                torch.ones((len(lengths),), dtype=torch.long),
                'center')
        else:
            raise ValueError(
                f'Interpolation method {method} is not defined')
        mask = mask_from_lengths(word_lengths)
        word_embeddings = word_embeddings.squeeze(2).transpose(0, 1).reshape(
            word_embeddings.shape[1],
            word_bounds_.shape[0],
            word_bounds_.shape[2]).permute(1, 0, 2) * mask
        inter['word_embeddings'] = word_embeddings
        word_embeddings = layers(
            state, 'word_decoder', word_embeddings, word_lengths, config)
    else:
        frame_embeddings = layers(                               # :92-94
            state, 'frame_encoder', input_layer(features), frame_lengths,
            config)
        inter['frame_embeddings'] = frame_embeddings
        if location == 'intermediate':                           # :96-107
            word_embeddings = downsample(
                frame_embeddings, word_bounds_, word_lengths, method)
            inter['word_embeddings'] = word_embeddings
            word_embeddings = layers(
                state, 'word_decoder', word_embeddings, word_lengths, config)
        elif location == 'loss':                                 # :109-115
            word_embeddings = downsample(
                frame_embeddings, word_bounds_, word_lengths, method)
            inter['word_embeddings'] = word_embeddings
        elif location == 'inference':                            # :117-130
            if training:
                return output_layer(frame_embeddings)
            word_embeddings = downsample(
                frame_embeddings, word_bounds_, word_lengths, method)
            inter['word_embeddings'] = word_embeddings
        else:
            raise ValueError(
                f'Downsample location {location} not recognized')
    logits = output_layer(word_embeddings)                       # :138
    if return_intermediates:
        return logits, inter
    return logits


def postprocess(logits, loss='bce'):
    """emphases.postprocess, core.py:335-342"""
    if loss == 'bce':
        return torch.sigmoid(logits)
    if loss == 'mse':
        return torch.clamp(logits, 0., 1.)
    return logits


def loss(
    scores, targets, word_lengths, loss_fn='bce', frame_lengths=None,
    word_bounds_=None, upsample_method=None):
    """emphases.loss, train/core.py:315-353.  With `upsample_method` set it is
    the training branch of the 'inference' location: targets are upsampled to
    frame resolution (and clamped for 'linear') and the mask covers frames."""
    if upsample_method is not None:
        targets = upsample(
            targets, word_bounds_, word_lengths, frame_lengths, upsample_method)
        if upsample_method == 'linear':
            targets = torch.clamp(targets, min=0., max=1.)
        mask = mask_from_lengths(frame_lengths)
    else:
        mask = mask_from_lengths(word_lengths)
    if loss_fn == 'bce':
        return torch.nn.functional.binary_cross_entropy_with_logits(
            scores[mask], targets[mask])
    if loss_fn == 'mse':
        return torch.nn.functional.mse_loss(scores[mask], targets[mask])
    raise ValueError(f'Loss {loss_fn} is not recognized')


###############################################################################
# Evaluation caller (emphases/evaluate/core.py:14-127, evaluate/metrics.py)
###############################################################################


def _welford(values):
    """torchutil.metrics.MeanStd (third party, absent: restated, PARITY
    UNPINNED): running mean and SAMPLE standard deviation over python floats,
    as `Statistics.update(values[mask].flatten().tolist())` feeds it
    (evaluate/metrics.py:103-111)"""
    count, mean, m2 = 0, 0., 0.
    for value in values:
        count += 1
        delta = value - mean
        mean += delta / count
        m2 += delta * (value - mean)
    return mean, math.sqrt(m2 / (count - 1))


def _bce_values(logits, targets, loss_fn):
    """evaluate/metrics.py:59-79"""
    if loss_fn == 'bce':
        return torch.nn.functional.binary_cross_entropy_with_logits(
            logits, targets, reduction='none')
    x = torch.clamp(logits, 0., 1.)
    return -(
        targets * torch.log(x + 1e-6) +
        (1 - targets) * torch.log(1 - x + 1e-6))


def evaluate(logits, targets, loss_fn='bce'):
    """The arithmetic of emphases.evaluate.datasets for one dataset.

    logits, targets: lists with one (W_i,) fp32 tensor per file (the
    un-postprocessed network output of evaluate/core.py:73-94 and the ground
    truth).  Pass 1 (core.py:27-46): mean / std of the postprocessed scores and
    of the targets over the whole dataset.  Pass 2 (core.py:56-110): per file
    and per dataset, Pearson correlation against those global statistics
    (torchutil PearsonCorrelation: sum((p - mean_p)(t - mean_t)) / (n std_p
    std_t)), mean BCE of the logits, mean squared error of the scores.
    Returns (overall dict, [per-file dict])."""
    scores = [postprocess(x, loss_fn) for x in logits]
    mean_p, std_p = _welford(torch.cat(scores).tolist())
    mean_t, std_t = _welford(torch.cat(targets).tolist())

    def metrics(xs, ps, ts):
        cross = sum(((p - mean_p) * (t - mean_t)).sum() for p, t in zip(ps, ts))
        count = sum(p.numel() for p in ps)
        bce = sum(_bce_values(x, t, loss_fn).sum() for x, t in zip(xs, ts))
        mse = sum(((p - t) ** 2).sum() for p, t in zip(ps, ts))
        return {
            'pearson_correlation':
                (1. / count * (cross / (std_p * std_t))).item(),
            'bce': (bce / count).item(),
            'mse': (mse / count).item()}

    granular = [
        metrics([x], [p], [t]) for x, p, t in zip(logits, scores, targets)]
    return metrics(logits, scores, targets), granular


###############################################################################
# Training batch sampler (emphases/data/sampler.py, dataset.py:97-117)
###############################################################################


def length_buckets(lengths, count=2):
    """Dataset.buckets, dataset.py:97-117"""
    size = len(lengths) // count
    indices = np.argsort(lengths)
    ordered = np.sort(lengths)
    parts = [
        np.stack((indices[i:i + size], ordered[i:i + size])).T
        for i in range(0, len(lengths), size)]
    if len(parts) == count + 1:
        residual = parts.pop()
        parts[-1] = np.concatenate((parts[-1], residual), axis=0)
    return parts


def epoch_batches(lengths, epoch, max_frames=75000, seed=0, count=2):
    """Sampler.batch, sampler.py:47-83"""
    generator = torch.Generator()
    generator.manual_seed(seed + epoch)
    batches = []
    for bucket in length_buckets(lengths, count):
        bucket = bucket[torch.randperm(len(bucket), generator=generator).tolist()]
        batch, longest = [], 0
        for index, length in bucket:
            longest = max(longest, length)
            if batch and (len(batch) + 1) * longest > max_frames:
                batches.append(batch)
                longest = length
                batch = [index]
            else:
                batch.append(index)
        if batch:
            batches.append(batch)
    return [
        batches[i]
        for i in torch.randperm(len(batches), generator=generator).tolist()]


###############################################################################
# End-to-end (emphases/core.py:223-265, 295-332)
###############################################################################


def from_alignment_and_audio(
    times,
    audio,
    state,
    config=None,
    batch_size=None,
    basis=None,
    autocast=False):
    """emphases.from_alignment_and_audio for METHOD == 'neural'.

    autocast=False: fp32 oracle (parity target).  autocast=True: the
    reference's own inference_context numerics (core.py:594-610: bf16
    autocast on CPU) -- used for the timed CPU baseline only.
    Returns scores (1, W_total).
    """
    config = dict(default_config(), **(config or {}))
    scores = []
    for features, bounds in preprocess(times, audio, batch_size, basis):
        frame_lengths = torch.tensor([features.shape[-1]], dtype=torch.long)
        word_lengths = torch.tensor([bounds.shape[-1]], dtype=torch.long)
        with torch.no_grad():
            if autocast:
                with torch.autocast('cpu'):
                    logits = model_forward(
                        state, features, frame_lengths, bounds, word_lengths,
                        config)
            else:
                logits = model_forward(
                    state, features, frame_lengths, bounds, word_lengths,
                    config)
        scores.append(postprocess(logits[0], config['LOSS']))
    return torch.cat(scores, 1)


###############################################################################
# Synthetic inputs (shared by tests and bench)
###############################################################################


def synthetic_alignment(duration, words, generator, min_frames=2):
    """`words` words tiling [0, duration] s with sorted uniform cuts; every
    word spans at least `min_frames` hops after int() truncation."""
    while True:
        cuts = torch.sort(
            torch.rand(words - 1, generator=generator, dtype=torch.float64)
            * duration).values.tolist()
        edges = [0.0] + cuts + [float(duration)]
        times = list(zip(edges[:-1], edges[1:]))
        bounds = word_bounds(times)
        if all(end - start >= min_frames for start, end in bounds):
            return times


def synthetic_utterance(seed, duration=None, words=None):
    """(times, audio (1, T) fp32) for one synthetic 16 kHz utterance"""
    generator = torch.Generator().manual_seed(seed)
    if duration is None:
        duration = 2.0 + 18.0 * float(torch.rand(1, generator=generator))
    samples = int(duration * SAMPLE_RATE) // HOPSIZE * HOPSIZE
    duration = samples / SAMPLE_RATE
    if words is None:
        words = max(2, int(2.5 * duration))
    audio = (0.1 * torch.randn(1, samples, generator=generator)).clamp(-1, 1)
    times = synthetic_alignment(duration, words, generator)
    return times, audio
