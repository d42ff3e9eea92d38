"""Data parallelism for the training step: shard the minibatch across ranks (independent utterances), one
all-reduce of the flat gradient bucket between `loss.backward()` and `optimizer.step()` (the reference is
single-process; this is where the collective goes in its loop, timit/steps/train_ctc.py:63 -> :65).

One process per GPU, `torch.distributed` for the plumbing (NCCL over NVLink/NVSwitch on the GPU box, gloo in
the CPU tests). BatchNorm statistics stay per-rank (DDP semantics): the G-rank gradient is the
mean of the G per-shard gradients because the reference divides the loss by the per-call batch size.
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise the default process group from RANK / WORLD_SIZE / MASTER_* (torchrun). Returns (rank, world)."""
    import datetime
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kwargs = {}
        if backend == "nccl":
            # bind this process to its GPU *before* the communicator exists; collectives then never guess a device
            local = int(os.environ.get("LOCAL_RANK", str(rank)))
            torch.cuda.set_device(local)
            kwargs["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world,
                                timeout=datetime.timedelta(seconds=int(os.environ.get("CTCB200_DIST_TIMEOUT", "300"))),
                                **kwargs)
    return rank, world


def shard_range(n_items, rank, world):
    """Contiguous utterance range [lo, hi) of `rank`: rank r of G takes utterances [r*N/G, (r+1)*N/G)."""
    lo = (n_items * rank) // world
    hi = (n_items * (rank + 1)) // world
    return lo, hi


class GradBucket(object):
    """Flat fp32 gradient bucket over a parameter list; `.allreduce_mean()` = one collective per step."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        self.numel = sum(p.numel() for p in self.params)
        p0 = self.params[0]
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=p0.device)
        self.views = []
        off = 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()

    def attach(self):
        """Make every p.grad a view into the bucket so backward accumulates straight into it."""
        self.flat.zero_()
        for p, v in zip(self.params, self.views):
            p.grad = v

    def allreduce_mean(self):
        if dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.mul_(1.0 / dist.get_world_size())

    def nbytes(self):
        return self.numel * 4
