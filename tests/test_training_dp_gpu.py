"""2-GPU data-parallel training step over NCCL (needs >= 2 GPUs; run with
`gpurun --gpus 2 -- python -m pytest tests/test_training_dp_gpu.py -m gpu`)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, tmpdir):
    os.environ.update(
        MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank),
        WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world,
                            device_id=torch.device('cuda', rank))
    try:
        import emphases_b200 as emphases
        from test_training_gpu import padded_batch
        emphases.reset_configuration()
        torch.manual_seed(0)                    # identical replicas
        model = emphases.Model().cuda()
        batch = padded_batch(seed=10 + rank)    # a different shard per rank
        batch = (batch[0].cuda(),) + batch[1:4] + (batch[4].cuda(),)
        model.train()
        scores = model(*batch[:4])
        emphases.loss(scores, batch[4], *batch[1:4], training=True).backward()
        local = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
        everyone = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(everyone, local)
        emphases.training.allreduce_gradients(model)
        reduced = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
        expected = torch.stack(everyone).mean(0)
        assert reduced.numel() == 250881
        assert torch.allclose(reduced, expected, rtol=1e-5, atol=1e-8)
        # one optimizer step keeps the replicas identical
        optimizer = torch.optim.Adam(model.parameters(), lr=1e-3)
        emphases.training.train_step(model, optimizer, batch)
        flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
        copies = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(copies, flat)
        assert torch.equal(copies[0], copies[1])
        with open(os.path.join(tmpdir, f'ok{rank}'), 'w') as stream:
            stream.write('ok')
    finally:
        dist.destroy_process_group()


def test_two_gpu_data_parallel_step(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    port = 32500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert all((tmp_path / f'ok{r}').exists() for r in range(2))
