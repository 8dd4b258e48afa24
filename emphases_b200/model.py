"""`emphases.Model` mirror (emphases/model/core.py:13-138).

The module holds exactly the reference's parameters (same `state_dict` keys,
SURVEY.md A.8) so reference checkpoints load unchanged; `forward` has the
reference's signature and return shape but runs the sm_100a kernels of
libemphases_b200.so on packed rows.  There is no PyTorch compute fallback.
"""
import functools

import numpy as np
import torch

import emphases_b200 as emphases
from . import _lib, engine


class Convolution(torch.nn.Sequential):
    """Parameter container mirroring emphases/model/layers/convolution.py"""

    def __init__(self, kernel_size):
        conv_fn = functools.partial(
            torch.nn.Conv1d, kernel_size=kernel_size, padding='same')
        layers = []
        channels = emphases.CHANNELS
        for _ in range(emphases.LAYERS):
            layers.extend((
                conv_fn(channels, channels),
                emphases.ACTIVATION_FUNCTION()))
            if emphases.DROPOUT is not None:
                layers.append(torch.nn.Dropout(emphases.DROPOUT))
        super().__init__(*layers)

    def forward(self, x, _):
        raise _lib.EmphasesB200Error(
            'submodules are parameter containers; call Model.forward')


def Layers(**kwargs):
    """emphases/model/layers/__init__.py:7-14"""
    if emphases.ARCHITECTURE == 'convolution':
        return Convolution(**kwargs)
    if emphases.ARCHITECTURE == 'transformer':
        from .transformer import Transformer
        return Transformer()
    raise ValueError(
        f'Network layer {emphases.ARCHITECTURE} is not defined')


def mask_from_lengths(lengths):
    """emphases/model/core.py:146-149"""
    x = torch.arange(
        lengths.max(), dtype=lengths.dtype, device=lengths.device)
    return (x.unsqueeze(0) < lengths.unsqueeze(1)).unsqueeze(1)


class Model(torch.nn.Module):

    def __init__(self):
        super().__init__()
        # Configuration is captured at construction, like the reference
        # module structure is
        self.architecture = emphases.ARCHITECTURE
        self.location = emphases.DOWNSAMPLE_LOCATION
        self.layers = emphases.LAYERS
        self.dropout = emphases.DROPOUT
        self.activation = emphases.ACTIVATION_FUNCTION.__name__
        if self.location not in ('input', 'intermediate', 'inference', 'loss'):
            raise ValueError(
                f'Downsample location {self.location} not recognized')
        self.input_layer = torch.nn.Conv1d(
            emphases.NUM_FEATURES,
            emphases.CHANNELS,
            kernel_size=emphases.ENCODER_KERNEL_SIZE,
            padding='same')
        self.frame_encoder = Layers(kernel_size=emphases.ENCODER_KERNEL_SIZE)
        if self.location in ['input', 'intermediate']:
            self.word_decoder = Layers(
                kernel_size=emphases.DECODER_KERNEL_SIZE)
        self.output_layer = torch.nn.Conv1d(
            emphases.CHANNELS,
            1,
            kernel_size=emphases.DECODER_KERNEL_SIZE,
            padding='same')
        self._packed = None
        self._packed_key = None

    # -- weights -----------------------------------------------------------

    def packed_weights(self):
        """Kernel-layout weights, re-packed only when parameters change"""
        # (walking the module tree costs more than a small launch: the
        # Parameter objects are stable, only their storage / version change)
        parameters = getattr(self, '_parameter_list', None)
        if parameters is None:
            parameters = list(self.parameters())
            object.__setattr__(self, '_parameter_list', parameters)
        device = parameters[0].device
        key = (device, tuple((p.data_ptr(), p._version) for p in parameters))
        if self._packed_key != key:
            if self.architecture == 'transformer':
                from . import transformer
                self._packed = transformer.pack_weights(self, device)
            else:
                self._packed = engine.pack_weights(
                    self.state_dict(), device, self.layers, self.activation,
                    self.dropout, hasattr(self, 'word_decoder'))
            self._packed_key = key
        return self._packed

    # -- forward -----------------------------------------------------------

    def forward(self, features, frame_lengths, word_bounds, word_lengths):
        """features (B, F, T) fp32 CUDA; word_bounds (B, 2, Wmax) int64;
        returns logits (B, 1, Wmax) fp32 (or (B, 1, T) for the 'inference'
        location in training mode), as emphases/model/core.py:39-138."""
        if features.device.type != 'cuda':
            raise _lib.EmphasesB200Error(
                'emphases_b200.Model runs on CUDA tensors only '
                '(there is no CPU fallback)')
        if self.input_layer.in_channels != emphases.NUM_MELS:
            emphases.require_mel_features_only()
        if self.training and self.dropout is not None:
            # the reference applies torch.nn.Dropout after every activation in
            # training mode (emphases/model/layers/convolution.py:29-30)
            raise NotImplementedError(
                'dropout is not implemented in the training-mode forward: set '
                'DROPOUT=None or call model.eval()')
        if torch.is_grad_enabled() and any(
            p.requires_grad for p in self.parameters()
        ) and self.training:
            from . import training
            return training.forward_with_grad(
                self, features, frame_lengths, word_bounds, word_lengths)
        with torch.cuda.device(features.device):
            return run_forward(
                self, features, frame_lengths, word_bounds, word_lengths)


def word_rows(word_bounds, word_lengths, device):
    """Packed word-row arrays for padded (B, 2, Wmax) bounds: every item gets
    Wmax slots; slots j >= word_lengths[i] are marked (-1, -1)"""
    batch, _, wmax = word_bounds.shape
    starts, total = engine.packed_starts([wmax] * batch)
    bounds = word_bounds.detach().to('cpu', torch.int64).numpy()
    lengths = word_lengths.detach().to('cpu', torch.int64).numpy()
    word_seq = np.full(total, -1, dtype=np.int32)
    word_lo = np.zeros(total, dtype=np.int32)
    word_hi = np.zeros(total, dtype=np.int32)
    for b in range(batch):
        s = int(starts[b])
        word_seq[s:s + wmax] = b
        word_lo[s:s + wmax] = bounds[b, 0]
        word_hi[s:s + wmax] = bounds[b, 1]
        word_lo[s + int(lengths[b]):s + wmax] = -1
        word_hi[s + int(lengths[b]):s + wmax] = -1
    blob = torch.from_numpy(np.concatenate([
        starts.astype(np.int32), np.full(batch, wmax, dtype=np.int32),
        word_seq, word_lo, word_hi])).to(device)
    views = {
        'word_row_start': blob[:batch],
        'n_words': blob[batch:2 * batch],
        'word_seq': blob[2 * batch:2 * batch + total],
        'word_lo': blob[2 * batch + total:2 * batch + 2 * total],
        'word_hi': blob[2 * batch + 2 * total:]}
    return views, starts, total, bounds, lengths


def run_forward(model, features, frame_lengths, word_bounds, word_lengths):
    eng = emphases.get_engine(features.device)
    weights = model.packed_weights()
    method = emphases.DOWNSAMPLE_METHOD
    if method not in _lib.POOL:
        raise ValueError(f'Interpolation method {method} is not defined')
    precision = emphases.precision_code()
    if model.architecture == 'transformer':
        from . import transformer
        return transformer.run_forward(
            model, eng, weights, features, frame_lengths, word_bounds,
            word_lengths)
    if model.location == 'input':
        from . import segments
        return segments.run_forward_input(
            model, eng, weights, features, word_bounds, word_lengths,
            method, precision)

    batch, channels, frames = features.shape
    device = features.device
    # Convolution ignores frame_lengths (convolution.py:36-37): every item
    # is a sequence of all `frames` columns, padding included
    starts, total = engine.packed_starts([frames] * batch)
    rows_meta = torch.from_numpy(np.concatenate([
        starts.astype(np.int32), np.full(batch, frames, dtype=np.int32)])
    ).to(device)
    row_start, n_rows = rows_meta[:batch], rows_meta[batch:]
    row_seq = eng.row_index(row_start, n_rows, batch, total)
    rows = torch.empty((total, channels), dtype=torch.float32, device=device)
    features = features.detach().to(torch.float32).contiguous()
    _lib.call(
        'emph_pack_rows', _lib.ptr(features), batch, channels, frames,
        _lib.ptr(row_start), _lib.ptr(n_rows), _lib.ptr(row_seq), total,
        _lib.ptr(rows), _lib.stream_ptr())
    frame_rows = eng.conv_stack(
        rows, row_seq, weights.frame,
        engine.frame_precision(precision, weights.frame))

    if model.location == 'inference' and model.training:
        # frame-resolution logits (model/core.py:119-122)
        logits, _ = eng.head(
            frame_rows, row_seq, weights, _lib.HEAD_LOGITS, want_scores=False)
        index = torch.from_numpy(
            (starts[:, None] + np.arange(frames)[None]).astype(np.int64)
        ).to(device)
        return logits[index][:, None, :]

    views, word_starts, total_words, bounds, lengths = word_rows(
        word_bounds, word_lengths, device)
    wmax = word_bounds.shape[2]
    valid = np.arange(wmax)[None] < lengths[:, None]
    engine.validate_bounds(
        np.stack([bounds[:, 0][valid], bounds[:, 1][valid]], axis=1),
        np.full(int(valid.sum()), frames), method)
    pooled = eng.pool(
        frame_rows, row_start, n_rows, views['word_seq'], views['word_lo'],
        views['word_hi'], method)
    word_row_seq = eng.row_index(
        views['word_row_start'], views['n_words'], batch, total_words)
    if model.location == 'intermediate':
        pooled = eng.conv_stack(
            pooled, word_row_seq, weights.word,
            engine.word_precision(precision, weights.word))
    logits, _ = eng.head(
        pooled, word_row_seq, weights, _lib.HEAD_LOGITS, want_scores=False)
    index = torch.from_numpy(
        (word_starts[:, None] + np.arange(wmax)[None]).astype(np.int64)
    ).to(device)
    return logits[index][:, None, :]
