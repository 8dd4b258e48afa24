"""Mirror of emphases/data/preprocess/core.py:71-125 (`from_audio`)"""
import emphases_b200 as emphases


def from_audio(audio, gpu=None):
    """Preprocess one audio chunk: (C, L) -> features (1, NUM_MELS, L // 160)

    Only the mel feature exists here (pitch / periodicity / loudness are
    disabled by default in the reference and need the `penn` network)."""
    emphases.require_mel_features_only()
    return emphases.data.preprocess.mels.from_audio(audio, gpu)[None]
