"""Sample-rate conversion (`emphases.resample`, emphases/core.py:613-619).

The reference delegates to torchaudio.transforms.Resample (windowed-sinc
interpolation, Hann window, lowpass_filter_width 6, rolloff 0.99).  The filter
bank below restates torchaudio's construction (torchaudio/functional/
functional.py `_get_sinc_resample_kernel`, a third-party dependency) in numpy
with the same dtypes; the convolution runs in csrc/resample.cu.
"""
import functools
import math

import numpy as np
import torch

from . import _lib

LOWPASS_FILTER_WIDTH = 6
ROLLOFF = 0.99


@functools.lru_cache(maxsize=16)
def filter_bank(orig_freq, new_freq):
    """(kernels float32 (new, 2 * width + orig), width, orig, new) after
    dividing both rates by their gcd"""
    gcd = math.gcd(int(orig_freq), int(new_freq))
    orig, new = int(orig_freq) // gcd, int(new_freq) // gcd
    base_freq = min(orig, new) * ROLLOFF
    width = math.ceil(LOWPASS_FILTER_WIDTH * orig / base_freq)
    idx = np.arange(-width, width + orig, dtype=np.float64)[None, :] / orig
    # torch: int64 arange / int -> float32 division, then promoted to float64
    phase = (np.arange(0, -new, -1).astype(np.float32) / np.float32(new)).astype(np.float64)
    t = phase[:, None] + idx
    t *= base_freq
    np.clip(t, -LOWPASS_FILTER_WIDTH, LOWPASS_FILTER_WIDTH, out=t)
    window = np.cos(t * math.pi / LOWPASS_FILTER_WIDTH / 2) ** 2
    t *= math.pi
    scale = base_freq / orig
    with np.errstate(invalid='ignore', divide='ignore'):
        kernels = np.where(t == 0, 1.0, np.sin(t) / t)
    kernels = kernels * window * scale
    return kernels.astype(np.float32), width, orig, new


_device_banks = {}
_fused_banks = {}


def device_bank(sample_rate, target_rate, device):
    """The filter bank on `device` in the form the fused log-mel front end
    takes (csrc/logmel.cu, emph_logmel_resampled_*): the kernels plus, per
    output phase, the range of taps that are not exactly zero (the Hann window
    of torchaudio's kernel is zero outside 6 zero crossings, so of the
    2 * width + orig taps only ~2 * width carry weight)"""
    key = (int(sample_rate), int(target_rate), torch.device(device))
    if key not in _fused_banks:
        kernels, width, orig, new = filter_bank(int(sample_rate), int(target_rate))
        nonzero = kernels != 0
        tap_lo = np.where(nonzero.any(1), nonzero.argmax(1), 0).astype(np.int32)
        tap_hi = np.where(
            nonzero.any(1), kernels.shape[1] - nonzero[:, ::-1].argmax(1), 0).astype(np.int32)
        _fused_banks[key] = {
            'filter': torch.from_numpy(kernels).to(device),
            'tap_lo': torch.from_numpy(tap_lo).to(device),
            'tap_hi': torch.from_numpy(tap_hi).to(device),
            'orig': orig, 'new': new, 'width': width}
    return _fused_banks[key]


def resampled_lengths(lengths, sample_rate, target_rate):
    """ceil(new * T / orig): the lengths torchaudio's Resample produces"""
    gcd = math.gcd(int(sample_rate), int(target_rate))
    orig, new = int(sample_rate) // gcd, int(target_rate) // gcd
    return -(-(new * np.asarray(lengths, dtype=np.int64)) // orig)


def resample(audio, sample_rate, target_rate, device):
    """(C, T) float tensor -> (C, ceil(T * target / source)) on `device`"""
    kernels, width, orig, new = filter_bank(int(sample_rate), int(target_rate))
    key = (int(sample_rate), int(target_rate), device)
    if key not in _device_banks:
        _device_banks[key] = torch.from_numpy(kernels).to(device)
    bank = _device_banks[key]
    audio = audio.detach().to(device, torch.float32).contiguous()
    channels, length = audio.shape
    target = int(math.ceil(new * length / orig))
    out = torch.empty((channels, target), dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        for channel in range(channels):
            _lib.call(
                'emph_resample_f32', _lib.ptr(audio[channel]), length,
                _lib.ptr(bank), orig, new, width, _lib.ptr(out[channel]), target,
                _lib.stream_ptr())
    return out


def resample_packed(packed, sample_rate, target_rate, device):
    """A PackedAudio of int16 PCM or fp32 samples at `sample_rate` (host; or
    a StreamedPack over a caller's list of tensors) -> a PackedAudio of fp32
    audio at `target_rate` resident on `device`: one upload and ONE kernel
    launch for the whole corpus (the reference resamples utterance by
    utterance, emphases/core.py:613-619)."""
    import numpy as np
    from . import scheduler
    if packed.buffer.dtype not in (torch.int16, torch.float32):
        raise ValueError('resample_packed expects int16 PCM or fp32 samples')
    if isinstance(packed, scheduler.StreamedPack):
        # a caller's list of tensors: packed (and narrowed when lossless) in
        # one go on the native pool
        everything = list(range(len(packed)))
        packed.start([everything])
        host = packed.launch_source(0, 0, len(packed) - 1)
        packed.finish()
        packed = scheduler.PackedAudio(host, packed.offsets, packed.lengths)
    if packed.ready is not None and len(packed):
        packed.ready(len(packed) - 1)
    kernels, width, orig, new = filter_bank(int(sample_rate), int(target_rate))
    key = (int(sample_rate), int(target_rate), device)
    if key not in _device_banks:
        _device_banks[key] = torch.from_numpy(kernels).to(device)
    bank = _device_banks[key]
    lengths = -(-(new * packed.lengths) // orig)            # ceil(new * T / orig)
    offsets, total = scheduler.PackedAudio.layout(lengths)
    with torch.cuda.device(device):
        source = packed.buffer.to(device, non_blocking=True)
        meta = torch.from_numpy(np.concatenate([
            packed.offsets, packed.lengths, offsets, lengths]).astype(np.int64)
        ).to(device)
        count = len(lengths)
        out = torch.empty(total, dtype=torch.float32, device=device)
        name = 'emph_resample_packed_i16' if source.dtype == torch.int16 \
            else 'emph_resample_packed_f32'
        _lib.call(
            name, _lib.ptr(source), _lib.ptr(meta[:count]),
            _lib.ptr(meta[count:2 * count]), _lib.ptr(meta[2 * count:3 * count]),
            _lib.ptr(meta[3 * count:]), count, _lib.ptr(bank), orig, new, width,
            _lib.ptr(out), total, _lib.stream_ptr())
        torch.cuda.current_stream(device).synchronize()     # `source` / staging may be reused
    return scheduler.PackedAudio(out, offsets, lengths)
