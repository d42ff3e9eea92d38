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


def _model_sync_worker(rank, world, port, out):
    """CTC_Model.backward with grad_sync set (per-layer buckets reduced from inside the backward pass), on the CPU emulation of
    the C ABI over gloo: the host logic of SURVEY.md §8(e) end to end."""
    from ctc_pytorch_b200.dist import GradSync
    from ctc_pytorch_b200.model import CTC_Model
    from ctc_pytorch_b200 import synth
    from tests.emu_lib import emulated
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    T, N, F, H, L, C = 10, 6, 40, 128, 2, 8
    torch.manual_seed(11)
    rnn_param = {"rnn_input_size": F, "rnn_hidden_size": H, "rnn_layers": L, "rnn_type": nn.LSTM, "bidirectional": True,
                 "batch_norm": True}
    m = CTC_Model(rnn_param=rnn_param, num_class=C, drop_out=0.0)
    m.precision, m.overlap_wgrad = "x3", False
    x, frac, tg, tl = synth.synthetic_batch(T, N, F, C, 3, 2)
    lo, hi = shard_range(N, rank, world)
    m.grad_sync = GradSync(weight=(hi - lo) * world / float(N))
    m.train()
    with emulated():
        o = m(x[lo:hi])
        il = (frac[lo:hi] * T).long()
        (nn.CTCLoss(reduction="sum")(o, tg[lo:hi], il, tl[lo:hi]) / (hi - lo)).backward()
    assert m.grad_sync.bytes == 4 * sum(p.numel() for p in m.parameters())   # L + 1 buckets covered every parameter
    if rank == 0:
        torch.save({k: p.grad.clone() for k, p in m.named_parameters()}, out)
    dist.destroy_process_group()


def test_model_backward_with_grad_sync_equals_mean_of_shard_oracle_gradients(tmp_path):
    from ctc_pytorch_b200 import synth
    from oracle import model_ref
    world = 2
    out = str(tmp_path / "model_sync.pt")
    mp.spawn(_model_sync_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    got = torch.load(out)
    T, N, F, H, L, C = 10, 6, 40, 128, 2, 8
    torch.manual_seed(11)
    ref = model_ref.RefAcousticModel(F, H, L, C, batch_norm=True)
    sd0 = {k: v.clone() for k, v in ref.state_dict().items()}
    x, frac, tg, tl = synth.synthetic_batch(T, N, F, C, 3, 2)
    want = {k: torch.zeros_like(p) for k, p in ref.named_parameters()}
    ref.train()
    for r in range(world):
        lo, hi = shard_range(N, r, world)
        ref.load_state_dict(sd0)
        ref.zero_grad()
        o = ref(x[lo:hi])
        il = (frac[lo:hi] * T).long()
        (nn.CTCLoss(reduction="sum")(o, tg[lo:hi], il, tl[lo:hi]) / (hi - lo)).backward()
        for k, p in ref.named_parameters():
            want[k] += p.grad * ((hi - lo) / float(N))
    for k in want:
        e = float((got[k].double() - want[k].double()).norm() / (want[k].double().norm() + 1e-30))
        assert e < 2e-4, (k, e)
