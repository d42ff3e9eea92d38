"""Drop-in for the `nn.CTCLoss(reduction='sum')` the reference builds at timit/steps/train_ctc.py:144
and calls at train_ctc.py:47 / back-propagates through at train_ctc.py:63.

Same constructor and call surface as torch.nn.CTCLoss (blank, reduction, zero_infinity;
log_probs [T,N,C], targets [N,S] padded or 1-D concatenated, input/target lengths), but forward and
backward run the hand-written warp-per-utterance alpha/beta kernels of libctcb200 (csrc/ctc.cu).
CUDA tensors only; there is no CPU path.
"""
import torch
import torch.nn as nn

from . import _lib


def _as_len_tensor(x, device):
    if torch.is_tensor(x):
        return x.to(device=device, dtype=torch.int64).contiguous()
    return torch.as_tensor(list(x), dtype=torch.int64, device=device)


def _pad_targets(targets, target_lengths, N):
    """warp-ctc style 1-D concatenated targets -> zero-padded [N, S_max] (my_863_corpus/steps/data_loader.py:194)."""
    lens = target_lengths.tolist()
    smax = max(lens) if lens else 0
    out = torch.zeros((N, max(smax, 1)), dtype=torch.int64, device=targets.device)
    off = 0
    for i, l in enumerate(lens):
        if l:
            out[i, :l] = targets[off:off + l]
        off += l
    return out


class _CTCLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, log_probs, targets, input_lengths, target_lengths, blank, max_target_len, zero_infinity):
        L = _lib.lib()
        T, N, C = log_probs.shape
        lp = log_probs.detach()
        if lp.dtype != torch.float32:
            lp = lp.float()
        lp = lp.contiguous()
        ws_floats = L.dll.ctcb200_ctc_workspace_floats(T, N, max_target_len)
        alpha_ws = torch.empty(ws_floats, dtype=torch.float32, device=lp.device)
        nll = torch.empty(N, dtype=torch.float32, device=lp.device)
        L.call("ctcb200_ctc_loss_fwd", _lib.ptr(lp), _lib.ptr(targets), targets.stride(0), _lib.ptr(input_lengths),
               _lib.ptr(target_lengths), T, N, C, max_target_len, blank, _lib.ptr(alpha_ws), _lib.ptr(nll),
               _lib.stream())
        ctx.save_for_backward(lp, targets, input_lengths, target_lengths, alpha_ws, nll)
        ctx.blank = blank
        ctx.max_target_len = max_target_len
        ctx.in_dtype = log_probs.dtype
        ctx.zero_infinity = zero_infinity
        return nll

    @staticmethod
    def backward(ctx, grad_nll):
        L = _lib.lib()
        lp, targets, input_lengths, target_lengths, alpha_ws, nll = ctx.saved_tensors
        T, N, C = lp.shape
        g = grad_nll.detach().to(torch.float32).contiguous()
        grad = torch.empty_like(lp)
        L.call("ctcb200_ctc_loss_bwd", _lib.ptr(lp), _lib.ptr(targets), targets.stride(0), _lib.ptr(input_lengths),
               _lib.ptr(target_lengths), T, N, C, ctx.max_target_len, ctx.blank, _lib.ptr(alpha_ws), _lib.ptr(nll),
               _lib.ptr(g), 1.0, _lib.ptr(grad), _lib.stream())
        if ctx.zero_infinity:
            grad = torch.where(torch.isinf(nll).view(1, N, 1), torch.zeros_like(grad), grad)
        if ctx.in_dtype != torch.float32:
            grad = grad.to(ctx.in_dtype)
        return grad, None, None, None, None, None, None


@_lib.on_tensor_device
def ctc_loss(log_probs, targets, input_lengths, target_lengths, blank=0, reduction="mean", zero_infinity=False):
    """Functional form; argument meaning identical to torch.nn.functional.ctc_loss."""
    _lib.require_cuda(log_probs)
    if log_probs.dim() != 3:
        raise RuntimeError("log_probs must be [T, N, C], got shape %s" % (tuple(log_probs.shape),))
    dev = log_probs.device
    T, N, C = log_probs.shape
    input_lengths = _as_len_tensor(input_lengths, dev)
    target_lengths = _as_len_tensor(target_lengths, dev)
    if input_lengths.numel() != N or target_lengths.numel() != N:
        raise RuntimeError("input_lengths / target_lengths must have batch size %d" % N)
    targets = targets.to(device=dev, dtype=torch.int64)
    if targets.dim() == 1:
        targets = _pad_targets(targets, target_lengths, N)
    elif targets.dim() != 2 or targets.shape[0] != N:
        raise RuntimeError("targets must be [N, S] or 1-D concatenated")
    if targets.shape[1] == 0:
        targets = torch.zeros((N, 1), dtype=torch.int64, device=dev)
    if targets.stride(1) != 1:
        targets = targets.contiguous()
    # S_max only sizes the per-lane state block; the padded width is an upper bound known without a sync
    max_target_len = int(targets.shape[1])
    nll = _CTCLossFn.apply(log_probs, targets, input_lengths, target_lengths, int(blank), max_target_len, bool(zero_infinity))
    if zero_infinity:
        nll = torch.where(torch.isinf(nll), torch.zeros_like(nll), nll)
    if reduction == "none":
        return nll
    if reduction == "sum":
        return nll.sum()
    if reduction == "mean":
        return (nll / target_lengths.clamp(min=1).to(nll.dtype)).mean()
    raise ValueError("%s is not a valid value for reduction" % reduction)


class CTCLoss(nn.Module):
    """Same signature as torch.nn.CTCLoss; the reference instantiates it with reduction='sum'."""

    def __init__(self, blank=0, reduction="mean", zero_infinity=False):
        super(CTCLoss, self).__init__()
        self.blank = blank
        self.reduction = reduction
        self.zero_infinity = zero_infinity

    def forward(self, log_probs, targets, input_lengths, target_lengths):
        return ctc_loss(log_probs, targets, input_lengths, target_lengths, self.blank, self.reduction,
                        self.zero_infinity)
