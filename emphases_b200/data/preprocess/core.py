"""Mirror of emphases/data/preprocess/core.py:71-125 (`from_audio`)"""
import emphases_b200 as emphases


def from_audio(audio, gpu=None):
    """Preprocess one audio chunk: (C, L) -> features (1, NUM_MELS, L // 160)

    Only the mel feature exists here (pitch / periodicity / loudness are
    disabled by default in the reference and need the `penn` network)."""
    if (
        emphases.PITCH_FEATURE or
        emphases.PERIODICITY_FEATURE or
        emphases.LOUDNESS_FEATURE
    ):
        raise NotImplementedError(
            'pitch / periodicity / loudness features are out of scope')
    return emphases.data.preprocess.mels.from_audio(audio, gpu)[None]
