"""Evaluation caller (emphases/evaluate/core.py:14-127) through the batched
path: the reference runs every test file twice, one file at a time (once for
the dataset statistics, once for the metrics); here the dataset goes through
the packed kernels ONCE, the logits stay on the device, and two launches of
`emph_word_metric_sums` give the statistics and every per-file metric.
"""
import json

import numpy as np
import torch

import emphases_b200 as emphases
from . import metrics as _metrics


def corpus(
    alignments, audios, targets, sample_rate=None, checkpoint=None,
    batch_size=None, gpu=None, stems=None, model=None
):
    """Metrics of one dataset held in memory.

    alignments / audios: as from_alignments_and_audio; targets: one tensor of
    per-word ground truth per file ((W_i,), (1, W_i) or (1, 1, W_i)).
    Returns (overall, granular): the dataset's dict of pearson_correlation /
    bce / mse and a {stem: dict} per file, like overall.json / granular.json
    (evaluate/core.py:109-110)."""
    from .. import core, scheduler
    sample_rate = emphases.SAMPLE_RATE if sample_rate is None else sample_rate
    if isinstance(gpu, (list, tuple)):
        gpu = gpu[0] if gpu else None
    device = emphases.resolve_device(gpu)
    if emphases.METHOD != 'neural':
        raise ValueError(f'Method {emphases.METHOD} is not defined')
    if model is None:
        model = core.load_model(checkpoint, device)
    with torch.cuda.device(device):
        logits = scheduler.run_on_device(
            model, alignments, audios, sample_rate, batch_size, device,
            to_cpu=False, output='logits')
        counts = [int(x.shape[-1]) for x in logits]
        flat_targets = [t.reshape(-1).float() for t in targets]
        for index, (count, target) in enumerate(zip(counts, flat_targets)):
            if target.numel() != count:
                raise ValueError(
                    f'file {index}: {target.numel()} targets for {count} words')
        flat = torch.cat([x.reshape(-1) for x in logits]).float()
        host = torch.cat(flat_targets) if flat_targets else torch.zeros(0)
        device_targets = host.to(device, non_blocking=True)

        # pass 1 (core.py:27-46): dataset statistics of scores and targets
        first = _metrics.word_sums(flat, device_targets, counts, 0)
        total = int(np.sum(counts))
        mean_p, mean_t = first[:, 0].sum() / total, first[:, 1].sum() / total
        # pass 2 (core.py:56-110): centred sums, BCE, squared error per file
        second = _metrics.word_sums(
            flat, device_targets, counts, 1, mean_p, mean_t)
    std_p = float(np.sqrt(second[:, 0].sum() / (total - 1)))
    std_t = float(np.sqrt(second[:, 1].sum() / (total - 1)))

    def result(sums, count):
        return {
            'pearson_correlation': float(
                1. / count * (sums[2] / (std_p * std_t))),
            'bce': float(sums[3] / count),
            'mse': float(sums[4] / count)}

    stems = list(range(len(counts))) if stems is None else stems
    granular = {
        stem: result(sums, count)
        for stem, sums, count in zip(stems, second, counts)}
    return result(second.sum(0), total), granular


def datasets(datasets, checkpoint=None, gpu=None):
    """evaluate/core.py:14-127: evaluate the 'test' partition of each dataset
    from the reference's cache layout (`CACHE_DIR/<name>/{alignment,audio,
    scores}/<stem>.{TextGrid,wav,pt}`, stems from `PARTITION_DIR/<name>.json`)
    and write `EVAL_DIR/<CONFIG>/{overall,granular}.json`."""
    overall, granular = {}, {}
    for dataset in datasets:
        cache = emphases.CACHE_DIR / dataset
        with open(emphases.PARTITION_DIR / f'{dataset}.json') as file:
            stems = json.load(file)['test']
        alignments = [
            emphases.Alignment(cache / 'alignment' / f'{stem}.TextGrid')
            for stem in stems]
        audios = [
            emphases.load.audio(cache / 'audio' / f'{stem}.wav')
            for stem in stems]
        targets = [
            torch.load(cache / 'scores' / f'{stem}.pt', weights_only=False)
            for stem in stems]
        overall[dataset], files = corpus(
            alignments, audios, targets, emphases.SAMPLE_RATE, checkpoint,
            None, gpu, stems)
        granular.update(
            {f'{dataset}/{stem}': value for stem, value in files.items()})
    directory = emphases.EVAL_DIR / emphases.CONFIG
    directory.mkdir(exist_ok=True, parents=True)
    with open(directory / 'overall.json', 'w') as file:
        json.dump(overall, file, indent=4)
    with open(directory / 'granular.json', 'w') as file:
        json.dump(granular, file, indent=4)
    return overall, granular
