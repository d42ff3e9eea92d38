"""Drop-in for the reference's variable-length (packed) acoustic model and its loss, the 863 path:
my_863_corpus/steps/model.py (SequenceWise :37-56, BatchRNN :69-86, CTC_RNN :89-141) fed by
`nn.utils.rnn.pack_padded_sequence` (my_863_corpus/steps/lstm_ctc.py:41) and trained with warp-ctc's `CTCLoss()`
(lstm_ctc.py:9,171: softmax inside the loss, 1-D concatenated int targets, summed over the batch).

Same constructor, `state_dict()` keys (`rnns.{l}.batch_norm.module.*`, `rnns.{l}.rnn.*`, `fc.module.0.*`, `fc.module.1.weight`)
and forward contract: PackedSequence in -> padded [T, N, num_class+1] out, raw activations in training mode, log-softmax per
frame in eval mode, ZERO rows (uniform log-probs in eval mode) for padded frames.

Packed semantics on dense tensors — alignment instead of masks: the sm_100a recurrent kernels always scan all T
rows of a padded batch, so the forward direction runs on the left-aligned batch and the reverse direction on a RIGHT-aligned
copy (frame k of utterance n at row T - len_n + k): its scan T-1 -> 0 meets the last valid frame first, with the zero initial
state, and the padding only afterwards. `ctcb200_realign_rows` moves activations / gradients between the two alignments and
zeroes the padding; BatchNorm statistics are column sums over the zero-padded tensor divided by the valid-frame count. BPTT
over padded frames is harmless by construction (dh = 0 and dc = 0 there give dG = 0). The design was validated against the
unmodified reference at the oracle level first (oracle/packed_ref.py, tests/test_oracle.py::test_alignment_design_*).
"""
from collections import OrderedDict

import torch
import torch.nn as nn

from . import _lib
from .loss import ctc_loss
from .model import _RnnStackFn, _call, _unwrap

__all__ = ["SequenceWise", "PackedBatchRNN", "CTC_RNN", "WarpCTCLoss"]


class SequenceWise(nn.Module):
    """Parameter container with the reference's key names (`.module.`); executed by CTC_RNN's fused CUDA path."""

    def __init__(self, module):
        super(SequenceWise, self).__init__()
        self.module = module

    def forward(self, x):
        raise RuntimeError("SequenceWise is executed by CTC_RNN's fused CUDA path; call the model, not the layer")


class PackedBatchRNN(nn.Module):
    """BatchNorm (over the valid frames) + bidirectional bias-free LSTM over a packed batch (model.py:69-86 of the 863 tree)."""

    def __init__(self, input_size, hidden_size, rnn_type=nn.LSTM, bidirectional=False, batch_norm=True, dropout=0.1):
        super(PackedBatchRNN, self).__init__()
        self.input_size = input_size
        self.hidden_size = hidden_size
        self.bidirectional = bidirectional
        self.batch_norm = SequenceWise(nn.BatchNorm1d(input_size)) if batch_norm else None
        # the reference passes `dropout` to a ONE-layer nn.LSTM, where torch applies none: kept as a constructor argument only
        self.rnn = rnn_type(input_size=input_size, hidden_size=hidden_size, bidirectional=bidirectional, bias=False)

    def forward(self, x):
        raise RuntimeError("PackedBatchRNN is executed by CTC_RNN's fused CUDA path; call the model, not the layer")


class CTC_RNN(nn.Module):
    def __init__(self, rnn_input_size=40, rnn_hidden_size=768, rnn_layers=5, rnn_type=nn.LSTM, bidirectional=True,
                 batch_norm=True, num_class=28, drop_out=0.1):
        super(CTC_RNN, self).__init__()
        self.rnn_input_size = rnn_input_size
        self.rnn_hidden_size = rnn_hidden_size
        self.rnn_layers = rnn_layers
        self.rnn_type = rnn_type
        self.num_class = num_class
        self.num_directions = 2 if bidirectional else 1
        self.name = "CTC_RNN"
        self._drop_out = drop_out
        self.batch_tile = 0
        self.overlap_wgrad = False
        self.precision = "bf16"
        self.grad_sync = None
        self.mask_source = None
        # attributes the shared autograd node reads
        self.rnn_param = {"rnn_input_size": rnn_input_size, "rnn_hidden_size": rnn_hidden_size, "rnn_layers": rnn_layers,
                          "rnn_type": rnn_type, "bidirectional": bidirectional, "batch_norm": batch_norm}

        rnns = [("0", PackedBatchRNN(rnn_input_size, rnn_hidden_size, rnn_type=rnn_type, bidirectional=bidirectional,
                                     batch_norm=False))]
        for i in range(rnn_layers - 1):
            rnns.append(("%d" % (i + 1), PackedBatchRNN(self.num_directions * rnn_hidden_size, rnn_hidden_size, rnn_type=rnn_type,
                                                        bidirectional=bidirectional, dropout=drop_out, batch_norm=batch_norm)))
        self.rnns = nn.Sequential(OrderedDict(rnns))
        if batch_norm:
            fc = nn.Sequential(nn.BatchNorm1d(self.num_directions * rnn_hidden_size),
                               nn.Linear(self.num_directions * rnn_hidden_size, num_class + 1, bias=False))
        else:
            fc = nn.Linear(self.num_directions * rnn_hidden_size, num_class + 1, bias=False)
        self.fc = SequenceWise(fc)

    def _params(self):
        plist = []
        for layer in self.rnns.children():
            bn = _unwrap(layer.batch_norm)
            if bn is not None:
                plist += [bn.weight, bn.bias]
            plist += [layer.rnn.weight_ih_l0, layer.rnn.weight_hh_l0, layer.rnn.weight_ih_l0_reverse,
                      layer.rnn.weight_hh_l0_reverse]
        fc = _unwrap(self.fc)
        if isinstance(fc, nn.Sequential):
            plist += [fc[0].weight, fc[0].bias, fc[1].weight]
        else:
            plist += [fc.weight]
        return plist

    def forward(self, x, lengths=None):
        """x: PackedSequence (as the reference feeds it), or a zero-padded time-major [T, N, F] tensor with `lengths`.
        Returns padded [T, N, num_class + 1]: activations in training mode, per-frame log-softmax in eval mode."""
        if isinstance(x, nn.utils.rnn.PackedSequence):
            x, lens = nn.utils.rnn.pad_packed_sequence(x, batch_first=False)     # layout only: [T, N, F], zero padding
            lengths = lens
        elif lengths is None:
            raise ValueError("CTC_RNN.forward takes a PackedSequence, or a padded [T, N, F] tensor together with lengths")
        _lib.require_cuda(x)
        if self.rnn_type not in (nn.LSTM, nn.GRU, nn.RNN) or self.num_directions != 2:
            raise RuntimeError("the packed B200 path implements bidirectional nn.LSTM / nn.GRU / nn.RNN layers")
        if next(self.parameters()).device != x.device:
            raise RuntimeError("model parameters and input must live on the same CUDA device")
        with torch.cuda.device(x.device):
            x = x.float().contiguous()
            T, N, F_ = x.shape
            lens_host = [int(v) for v in (lengths.tolist() if torch.is_tensor(lengths) else lengths)]
            if len(lens_host) != N or max(lens_host) > T or min(lens_host) < 1:
                raise ValueError("lengths must hold one value in [1, T] per utterance")
            lens_dev = torch.as_tensor(lens_host, dtype=torch.int64, device=x.device)
            need_grad = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters()))
            geom = (T, N, F_, N * F_, F_, need_grad, lens_dev, self.training, sum(lens_host))
            return _RnnStackFn.apply(self, x, geom, *self._params())

    @staticmethod
    def save_package(model, optimizer=None, decoder=None, epoch=None, loss_results=None, training_cer_results=None,
                     dev_cer_results=None):
        package = {"input_size": model.rnn_input_size, "hidden_size": model.rnn_hidden_size, "rnn_layers": model.rnn_layers,
                   "rnn_type": model.rnn_type, "num_class": model.num_class, "_drop_out": model._drop_out,
                   "state_dict": model.state_dict()}
        if optimizer is not None:
            package["optim_dict"] = optimizer.state_dict()
        if decoder is not None:
            package["decoder"] = decoder
        if epoch is not None:
            package["epoch"] = epoch
        if loss_results is not None:
            package["loss_results"] = loss_results
            package["training_cer_results"] = training_cer_results
            package["dev_cer_results"] = dev_cer_results
        return package


class _LogSoftmaxFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, act):
        T, N, C = act.shape
        a = act.detach().float().contiguous()
        out = torch.empty_like(a)
        _call("ctcb200_log_softmax_fwd", _lib.ptr(a), C, _lib.ptr(out), T * N, C, _lib.stream())
        ctx.save_for_backward(out)
        return out

    @staticmethod
    def backward(ctx, g):
        (out,) = ctx.saved_tensors
        T, N, C = out.shape
        g = g.detach().float().contiguous()
        dx = torch.empty_like(out)
        _call("ctcb200_log_softmax_bwd", _lib.ptr(g), _lib.ptr(out), _lib.ptr(dx), T * N, C, _lib.stream())
        return dx


class WarpCTCLoss(nn.Module):
    """warp-ctc's `CTCLoss()` call surface (lstm_ctc.py:44,171; size_average=False): `loss_fn(activations [T,N,C],
    targets 1-D int (concatenated), input_sizes [N], target_sizes [N])` -> summed negative log likelihood of
    softmax(activations); blank = 0. Runs log-softmax and the alpha/beta kernels of libctcb200."""

    def __init__(self, size_average=False, blank=0):
        super(WarpCTCLoss, self).__init__()
        self.size_average = size_average
        self.blank = blank

    @_lib.on_tensor_device
    def forward(self, acts, labels, act_lens, label_lens):
        _lib.require_cuda(acts)
        lp = _LogSoftmaxFn.apply(acts)
        loss = ctc_loss(lp, labels.to(acts.device), act_lens, label_lens, blank=self.blank, reduction="sum")
        if self.size_average:
            loss = loss / acts.shape[1]
        return loss
