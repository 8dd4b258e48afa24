"""Audio loading, mirroring emphases/load.py:11-17 (load + resample to 16 kHz)

torchaudio.load needs torchcodec (absent in this image), so RIFF/WAVE PCM
files are decoded here; other containers fall back to torchaudio when it
works.
"""
import struct

import numpy as np
import torch

import emphases_b200 as emphases


def audio(file):
    """Load audio and maybe resample; returns (channels, samples) float32"""
    samples, sample_rate = wav(file)
    return emphases.resample(samples, sample_rate)


def wav(file, normalize=True):
    """(channels, samples) tensor and the sample rate of a wav file.
    normalize=False keeps int16 PCM as int16 (for the int16 upload path)."""
    with open(file, 'rb') as stream:
        raw = stream.read()
    if raw[:4] != b'RIFF' or raw[8:12] != b'WAVE':
        import torchaudio
        return torchaudio.load(file)
    cursor = 12
    fmt = None
    data = None
    while cursor + 8 <= len(raw):
        tag = raw[cursor:cursor + 4]
        size = struct.unpack('<I', raw[cursor + 4:cursor + 8])[0]
        body = raw[cursor + 8:cursor + 8 + size]
        if tag == b'fmt ':
            fmt = struct.unpack('<HHIIHH', body[:16])
            if fmt[0] == 0xFFFE and len(body) >= 26:       # extensible
                fmt = (struct.unpack('<H', body[24:26])[0],) + fmt[1:]
        elif tag == b'data':
            data = body
        cursor += 8 + size + (size & 1)
    if fmt is None or data is None:
        raise ValueError(f'{file}: not a PCM wav file')
    code, channels, sample_rate, _, _, bits = fmt
    if code == 1 and bits == 16:
        array = np.frombuffer(data, dtype='<i2')
        if not normalize:
            return torch.from_numpy(
                array.reshape(-1, channels).T.copy()), sample_rate
        array = array.astype(np.float32) / 32768.
    elif code == 1 and bits == 32:
        array = np.frombuffer(data, dtype='<i4').astype(np.float32) / 2147483648.
    elif code == 1 and bits == 8:
        array = (np.frombuffer(data, dtype=np.uint8).astype(np.float32) - 128.) / 128.
    elif code == 3 and bits == 32:
        array = np.frombuffer(data, dtype='<f4').astype(np.float32)
    else:
        raise ValueError(f'{file}: unsupported wav encoding {code}/{bits}')
    return torch.from_numpy(array.reshape(-1, channels).T.copy()), sample_rate


def save_wav(file, samples, sample_rate=16000):
    """Write (channels, samples) float tensor as 16-bit PCM"""
    pcm = (samples.clamp(-1, 1) * 32767.).round().to(torch.int16)
    body = pcm.T.contiguous().numpy().astype('<i2').tobytes()
    channels = pcm.shape[0]
    header = b'RIFF' + struct.pack('<I', 36 + len(body)) + b'WAVEfmt ' + \
        struct.pack(
            '<IHHIIHH', 16, 1, channels, sample_rate,
            sample_rate * channels * 2, channels * 2, 16) + \
        b'data' + struct.pack('<I', len(body))
    with open(file, 'wb') as stream:
        stream.write(header + body)
