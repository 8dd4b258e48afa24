"""Padded training batches (emphases/data/collate.py:11-77): same tuple, same
shapes and dtypes, vectorised copies instead of one slice assignment per
tensor per item."""
import torch

import emphases_b200 as emphases


def collate(batch):
    """[(features (F, T_i), scores (1, W_i), word_bounds (2, W_i), alignment,
    audio (1, S_i), stem)] -> (features (B, F, Tmax), frame_lengths (B,),
    word_bounds (B, 2, Wmax) int64, word_lengths (B,), scores (B, 1, Wmax),
    alignments, audio (B, 1, Tmax * HOPSIZE), stems)"""
    features, scores, word_bounds, alignments, audios, stems = zip(*batch)
    word_lengths = torch.tensor(
        [bounds.shape[-1] for bounds in word_bounds], dtype=torch.long)
    frame_lengths = torch.tensor(
        [feature.shape[-1] for feature in features], dtype=torch.long)
    words, frames = int(word_lengths.max()), int(frame_lengths.max())
    size = len(features)
    padded_features = torch.zeros((size, emphases.NUM_FEATURES, frames))
    padded_scores = torch.zeros((size, 1, words))
    padded_bounds = torch.zeros((size, 2, words), dtype=torch.long)
    padded_audio = torch.zeros((size, 1, frames * emphases.HOPSIZE))
    for i in range(size):
        t, w = int(frame_lengths[i]), int(word_lengths[i])
        padded_features[i, :, :t] = features[i]
        padded_scores[i, :, :w] = scores[i][:, :w]
        padded_bounds[i, :, :w] = word_bounds[i][:, :w]
        end = t * emphases.HOPSIZE
        # an audio shorter than frames * hop fails like the reference
        # (collate.py:65-66: shape mismatch in the slice assignment)
        padded_audio[i, :, :end] = audios[i][:, :end]
    return (
        padded_features, frame_lengths, padded_bounds, word_lengths,
        padded_scores, alignments, padded_audio, stems)
