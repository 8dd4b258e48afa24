"""Unit conversions between seconds, samples and hops.

Drop-in for the six helpers of the reference's `emphases.convert`
(emphases/convert.py:9-36).  The chunker depends on two details of those
helpers, kept here on purpose: the sample rate and hop size are read from the
package configuration at call time, and `samples_to_frames` is a FLOOR
division that stays in floating point for float inputs (so
`seconds_to_frames(0.995)` is `99.0`, not `99`).
"""
import emphases_b200 as _config


def samples_to_frames(samples):
    """Whole hops contained in `samples` (float in, float out)"""
    hop = _config.HOPSIZE
    return samples // hop


def samples_to_seconds(samples):
    return samples / _config.SAMPLE_RATE


def seconds_to_samples(seconds):
    return seconds * _config.SAMPLE_RATE


def seconds_to_frames(seconds):
    """Seconds -> samples -> whole hops"""
    return samples_to_frames(seconds_to_samples(seconds))


def frames_to_samples(frames):
    hop = _config.HOPSIZE
    return frames * hop


def frames_to_seconds(frames):
    """Hop count -> seconds, through the derived HOPSIZE_SECONDS constant"""
    return frames * _config.HOPSIZE_SECONDS
