"""One-process-per-GPU execution of a corpus (torchrun / torch.distributed).

Utterances are independent (emphases/core.py:169-179), so ranks share nothing
on the data path: every rank takes an LPT-balanced shard of the file list,
runs it on its own GPU and writes its own `{prefix}.TextGrid` / `{prefix}.pt`
files.  The process group (NCCL on GPUs, gloo in CPU tests) is used only to
agree on the shards and, optionally, to gather scores on rank 0.
"""
import os

import torch

import emphases_b200 as emphases
from . import scheduler


def rank_and_world():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))


def shard(costs, rank=None, world=None):
    """Indices of this rank's utterances: deterministic length-balanced split
    by cost (frames; frames**2 for the transformer variant) -- identical on
    every rank, so no communication is needed to agree on it.  LPT for short
    lists, its vectorised snake form for long ones."""
    if rank is None or world is None:
        rank, world = rank_and_world()
    if world == 1:
        return list(range(len(costs)))
    if len(costs) > 2048:
        return scheduler.snake_assign(costs, world)[rank]
    return scheduler.lpt_assign(list(costs), world)[rank]


def audio_costs(audio_files):
    """Cost proxy without decoding: file size in bytes ~ samples (native stat
    on the worker pool: every rank looks at the whole list)"""
    import ctypes
    import numpy as np
    from . import _lib, corpus
    files = [os.fspath(file) for file in audio_files]
    sizes = np.zeros(max(len(files), 1), dtype=np.int64)
    local_world = max(1, int(os.environ.get('LOCAL_WORLD_SIZE', '1')))
    threads = max(2, min(32, (os.cpu_count() or 1) // local_world))
    _lib.load().emph_file_sizes(
        corpus._blob(files), len(files), threads, sizes.ctypes.data_as(ctypes.c_void_p))
    sizes = sizes[:len(files)]
    if np.any(sizes < 0):
        raise FileNotFoundError(files[int(np.nonzero(sizes < 0)[0][0])])
    return sizes


def from_files_to_files(
    text_files,
    audio_files,
    output_prefixes=None,
    checkpoint=None,
    batch_size=None,
    gpu=None
):
    """emphases.from_files_to_files across the ranks of a process group:
    each rank annotates its shard on GPU `gpu` (default LOCAL_RANK)."""
    rank, world = rank_and_world()
    if output_prefixes is None:
        output_prefixes = [os.path.splitext(str(f))[0] for f in text_files]
    mine = list(range(len(audio_files))) if world == 1 else \
        shard(audio_costs(audio_files), rank, world)
    if gpu is None:
        gpu = int(os.environ.get('LOCAL_RANK', 0))
    if mine:
        emphases.from_files_to_files(
            [text_files[i] for i in mine],
            [audio_files[i] for i in mine],
            [output_prefixes[i] for i in mine],
            checkpoint, batch_size, gpu)
    return mine


def gather_scores(indices, scores, total, destination=0):
    """Gather per-utterance score tensors on `destination` (host-side gather
    of small tensors; returns the full list there, None elsewhere)."""
    import torch.distributed as dist
    rank, world = rank_and_world()
    payload = [(int(i), s.cpu()) for i, s in zip(indices, scores)]
    if world == 1:
        gathered = [payload]
    else:
        gathered = [None] * world if rank == destination else None
        dist.gather_object(payload, gathered, dst=destination)
    if rank != destination:
        return None
    result = [None] * total
    for part in gathered:
        for index, score in part:
            result[index] = score
    return result


def reduce_timing(elapsed_ms, units, device=None):
    """Max time over ranks and summed units: the multi-GPU throughput rule"""
    import torch.distributed as dist
    rank, world = rank_and_world()
    stats = torch.tensor([elapsed_ms, units], dtype=torch.float64, device=device)
    if world > 1:
        worst, total = stats.clone(), stats.clone()
        dist.all_reduce(worst, op=dist.ReduceOp.MAX)
        dist.all_reduce(total, op=dist.ReduceOp.SUM)
        return worst[0].item(), total[1].item()
    return float(elapsed_ms), float(units)
