"""CPU suite: the N>1 host logic (minibatch sharding + one flat-bucket gradient all-reduce) on world_size-2 gloo."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

from ctc_pytorch_b200.dist import GradBucket, shard_range


def test_shard_range_partitions_the_batch():
    for n in (1, 7, 32, 64, 65):
        for world in (1, 2, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    model = nn.Sequential(nn.Linear(6, 5), nn.Linear(5, 3))   # same weights on every rank
    bucket = GradBucket(model.parameters())
    g = torch.Generator().manual_seed(100)
    x = torch.randn(8, 6, generator=g)
    lo, hi = shard_range(8, rank, world)
    bucket.attach()
    # the reference divides the summed loss by the per-call batch size (train_ctc.py:48)
    loss = model(x[lo:hi]).pow(2).sum() / (hi - lo)
    loss.backward()
    bucket.allreduce_mean()
    if rank == 0:
        torch.save([p.grad.clone() for p in model.parameters()], out)
    dist.destroy_process_group()


def test_two_rank_allreduce_equals_mean_of_shard_gradients(tmp_path):
    world = 2
    out = str(tmp_path / "grads.pt")
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    got = torch.load(out)
    torch.manual_seed(0)
    model = nn.Sequential(nn.Linear(6, 5), nn.Linear(5, 3))
    g = torch.Generator().manual_seed(100)
    x = torch.randn(8, 6, generator=g)
    want = [torch.zeros_like(p) for p in model.parameters()]
    for r in range(world):
        lo, hi = shard_range(8, r, world)
        model.zero_grad()
        (model(x[lo:hi]).pow(2).sum() / (hi - lo)).backward()
        for w, p in zip(want, model.parameters()):
            w += p.grad / world
    for a, b in zip(got, want):
        assert torch.allclose(a, b, atol=1e-6)
    # p.grad are views into one flat bucket: a single collective covers every parameter
    bucket = GradBucket(model.parameters())
    bucket.attach()
    assert all(p.grad.data_ptr() >= bucket.flat.data_ptr() for p in model.parameters())
    assert bucket.nbytes() == 4 * sum(p.numel() for p in model.parameters())
