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


def _sync_worker(rank, world, port, out, n_total):
    """GradSync as CTC_Model.backward drives it: one reduce() per layer bucket in reverse layer order, then wait()."""
    from ctc_pytorch_b200.dist import GradSync
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    model = nn.Sequential(nn.Linear(6, 5), nn.Linear(5, 3))
    g = torch.Generator().manual_seed(100)
    x = torch.randn(n_total, 6, generator=g)
    lo, hi = shard_range(n_total, rank, world)
    loss = model(x[lo:hi]).pow(2).sum() / (hi - lo)        # per-call batch size, train_ctc.py:48
    loss.backward()
    sync = GradSync(weight=(hi - lo) * world / float(n_total))
    buckets = []
    for layer in reversed(list(model.children())):       # "per layer, last layer first"
        flat = torch.cat([p.grad.reshape(-1) for p in layer.parameters()])
        sync.reduce(flat)
        buckets.append(flat)
    sync.wait()
    assert sync.bytes == 4 * sum(p.numel() for p in model.parameters())
    if rank == 0:
        torch.save(buckets, out)
    dist.destroy_process_group()


def test_grad_sync_weighted_mean_equals_full_batch_gradient(tmp_path):
    """Uneven shards (7 utterances over 2 ranks): the shard-size-weighted mean of the rank gradients is the gradient of the
    full-batch mean loss (no cross-utterance coupling in this toy model), which a plain mean would get wrong."""
    world, n_total = 2, 7
    out = str(tmp_path / "sync.pt")
    mp.spawn(_sync_worker, args=(world, _free_port(), out, n_total), nprocs=world, join=True)
    got = torch.load(out)
    torch.manual_seed(0)
    model = nn.Sequential(nn.Linear(6, 5), nn.Linear(5, 3))
    g = torch.Generator().manual_seed(100)
    x = torch.randn(n_total, 6, generator=g)
    (model(x).pow(2).sum() / n_total).backward()
    want = [torch.cat([p.grad.reshape(-1) for p in layer.parameters()]) for layer in reversed(list(model.children()))]
    for a, b in zip(got, want):
        assert torch.allclose(a, b, atol=1e-6)


def test_grad_bucket_detects_severed_views():
    import pytest
    model = nn.Linear(3, 2)
    bucket = GradBucket(model.parameters())
    bucket.attach()
    model(torch.randn(4, 3)).sum().backward()
    bucket.allreduce_mean()                  # single process: nothing to reduce, views intact
    for p in model.parameters():
        p.grad = None                        # what optimizer.zero_grad(set_to_none=True) does
    model(torch.randn(4, 3)).sum().backward()
    with pytest.raises(RuntimeError, match="attach"):
        bucket.allreduce_mean()
