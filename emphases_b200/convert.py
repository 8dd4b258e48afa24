"""Time conversions, mirroring emphases/convert.py:9-36"""
import emphases_b200 as emphases


def frames_to_samples(frames):
    return frames * emphases.HOPSIZE


def frames_to_seconds(frames):
    return frames * emphases.HOPSIZE_SECONDS


def seconds_to_frames(seconds):
    return samples_to_frames(seconds_to_samples(seconds))


def seconds_to_samples(seconds):
    return seconds * emphases.SAMPLE_RATE


def samples_to_frames(samples):
    # float floor division on float inputs, like the reference
    return samples // emphases.HOPSIZE


def samples_to_seconds(samples):
    return samples / emphases.SAMPLE_RATE
