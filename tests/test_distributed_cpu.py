"""world_size-2 gloo tests of the multi-process (one process per GPU) path"""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, tmpdir):
    os.environ.update(
        MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank),
        WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from emphases_b200 import distributed
        costs = [9, 1, 7, 3, 5, 5, 2, 8, 4, 6, 1]
        mine = distributed.shard(costs)
        # every rank computes the same split: gather and compare
        everyone = [None] * world
        dist.all_gather_object(everyone, mine)
        assert sorted(i for part in everyone for i in part) == list(range(len(costs)))
        loads = [sum(costs[i] for i in part) for part in everyone]
        assert max(loads) - min(loads) <= max(costs) // 2
        # per-utterance results come back in corpus order on rank 0
        scores = [torch.full((1, 3), float(i)) for i in mine]
        gathered = distributed.gather_scores(mine, scores, len(costs))
        if rank == 0:
            assert [int(s[0, 0]) for s in gathered] == list(range(len(costs)))
        else:
            assert gathered is None
        # throughput rule: max time over ranks, summed units
        elapsed, units = distributed.reduce_timing(10.0 * (rank + 1), 100.0)
        assert elapsed == 10.0 * world and units == 100.0 * world
        with open(os.path.join(tmpdir, f'ok{rank}'), 'w') as stream:
            stream.write('ok')
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_and_gather(tmp_path):
    world = 2
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f'ok{r}').exists() for r in range(world))


def _grad_worker(rank, world, port, tmpdir):
    os.environ.update(
        MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank),
        WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from emphases_b200 import training
        torch.manual_seed(0)
        model = torch.nn.Sequential(
            torch.nn.Conv1d(4, 4, 3), torch.nn.ReLU(), torch.nn.Conv1d(4, 1, 3))
        for index, parameter in enumerate(model.parameters()):
            parameter.grad = torch.full_like(parameter, float(rank + 1) * (index + 1))
        training.allreduce_gradients(model)
        for index, parameter in enumerate(model.parameters()):
            expected = (index + 1) * sum(range(1, world + 1)) / world
            assert torch.allclose(parameter.grad, torch.full_like(parameter, expected))
        with open(os.path.join(tmpdir, f'grad{rank}'), 'w') as stream:
            stream.write('ok')
    finally:
        dist.destroy_process_group()


def test_flat_bucket_gradient_allreduce(tmp_path):
    world = 2
    port = 31500 + os.getpid() % 2000
    mp.spawn(_grad_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f'grad{r}').exists() for r in range(world))
