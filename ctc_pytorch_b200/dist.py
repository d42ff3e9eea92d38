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


class GradSync(object):
    """Gradient all-reduce launched from INSIDE the backward pass, one collective per layer bucket.

    CTC_Model.backward hands every layer's gradients (one flat fp32 buffer: W_ih, W_hh of both directions, the BatchNorm
    affine) to `reduce()` as soon as they are complete; the collective runs asynchronously (NCCL stream) behind the BPTT
    kernels of the layers still to be back-propagated, and `wait()` joins everything before autograd hands the gradients to
    the optimizer. Replaces "one flat all-reduce after backward()" (exposed: 84.8 MB at cfg2) by L+1 overlapped ones.

    The result is the shard-size-weighted mean: the reference divides the loss by the per-call batch size
    (timit/steps/train_ctc.py:48), so rank r's gradient is that of its own mean loss and the full-batch gradient is
    sum_r (n_r / n_total) g_r. `weight` = n_r * world / n_total (1.0 for equal shards); BatchNorm statistics stay per rank
    (DDP semantics)."""

    def __init__(self, weight=1.0, group=None):
        self.weight = float(weight)
        self.group = group
        self.pending = []
        self.bytes = 0

    def reduce(self, flat):
        if not (dist.is_initialized() and dist.get_world_size(self.group) > 1):
            return
        world = dist.get_world_size(self.group)
        flat.mul_(self.weight / world)           # pre-scaled: the SUM below is the weighted mean on every backend
        self.pending.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        self.bytes += flat.numel() * flat.element_size()

    def wait(self):
        for h in self.pending:
            h.wait()     # NCCL: makes the current stream wait for the collective; gloo: blocks the host
        self.pending = []


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
        for p, v in zip(self.params, self.views):
            # optimizer.zero_grad(set_to_none=True) severs the views: reducing a stale zero buffer would silently
            # leave the ranks unsynchronised
            if p.grad is None or p.grad.data_ptr() != v.data_ptr():
                raise RuntimeError("GradBucket: p.grad is no longer a view of the bucket; call attach() before backward()")
        if dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.mul_(1.0 / dist.get_world_size())

    def nbytes(self):
        return self.numel * 4
