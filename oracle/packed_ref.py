"""TEST INFRASTRUCTURE ONLY — variable-length ("packed") semantics of the 863 path, restated without
PackedSequence: my_863_corpus/steps/model.py:37-141 (SequenceWise, BatchRNN, CTC_RNN) fed by pack_padded_sequence
(lstm_ctc.py:41) with a warp-ctc style loss on raw activations (lstm_ctc.py:171, 1-D concatenated int targets,
data_loader.py:168-198).

What the reference does, stated on padded [T, N, .] tensors and per-utterance lengths len_n:
  * BatchNorm1d of layers >= 1 and of the output layer: statistics over the VALID frames only (rows (t, n) with
    t < len_n, sum(len) of them), applied to valid frames;
  * nn.LSTM(bias=False, bidirectional): the forward direction runs t = 0 .. len_n-1, the reverse direction starts from a
    zero state at t = len_n-1 and runs down to 0; nothing is computed for t >= len_n;
  * Linear(2H -> C+1, no bias) on valid frames; after pad_packed_sequence every padded frame is a ZERO vector;
  * training returns raw activations (the loss applies softmax itself), inference returns log_softmax per frame (zero
    rows therefore become uniform log-probabilities);
  * loss = sum over utterances of the CTC negative log likelihood of softmax(activations) over the first len_n frames
    (warp-ctc, size_average=False), targets concatenated into one 1-D int tensor with target_sizes.
Each utterance is pushed through torch's own LSTM on its valid slice, which is an independent formulation from packing.
"""
from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F


def _masked_bn(bn, x, lengths):
    """BatchNorm1d `bn` over the valid rows of x [T, N, C]; padded rows come back as zeros."""
    T, N, C = x.shape
    mask = (torch.arange(T).unsqueeze(1) < torch.as_tensor(lengths).unsqueeze(0))       # [T, N]
    rows = x[mask]                                                                       # [sum(len), C]
    out = torch.zeros_like(x)
    out[mask] = bn(rows)
    return out


class RefPackedModel(nn.Module):
    """Same parameter names / shapes as the reference's CTC_RNN (so state_dicts interchange)."""

    def __init__(self, rnn_input_size=40, rnn_hidden_size=128, rnn_layers=2, batch_norm=True, num_class=28):
        super().__init__()
        self.hidden, self.layers, self.num_class = rnn_hidden_size, rnn_layers, num_class

        class _Layer(nn.Module):
            def __init__(self, inp, hid, bn):
                super().__init__()
                self.batch_norm = _SeqWise(nn.BatchNorm1d(inp)) if bn else None
                self.rnn = nn.LSTM(input_size=inp, hidden_size=hid, bidirectional=True, bias=False)

        class _SeqWise(nn.Module):          # keeps the reference's key names: rnns.1.batch_norm.module.weight ...
            def __init__(self, module):
                super().__init__()
                self.module = module

        blocks = [("0", _Layer(rnn_input_size, rnn_hidden_size, False))]
        for i in range(rnn_layers - 1):
            blocks.append(("%d" % (i + 1), _Layer(2 * rnn_hidden_size, rnn_hidden_size, batch_norm)))
        self.rnns = nn.Sequential(OrderedDict(blocks))
        if batch_norm:
            fc = nn.Sequential(nn.BatchNorm1d(2 * rnn_hidden_size), nn.Linear(2 * rnn_hidden_size, num_class + 1, bias=False))
        else:
            fc = nn.Linear(2 * rnn_hidden_size, num_class + 1, bias=False)
        self.fc = _SeqWise(fc)

    def forward(self, x, lengths):
        """x [T, N, F] zero-padded, lengths [N] (any order) -> activations (train) / log-probs (eval) [T, N, C+1]."""
        T, N, _ = x.shape
        lengths = [int(v) for v in lengths]
        for layer in self.rnns.children():
            if layer.batch_norm is not None:
                x = _masked_bn(layer.batch_norm.module, x, lengths)
            out = torch.zeros(T, N, 2 * self.hidden, dtype=x.dtype)
            for n in range(N):
                y, _ = layer.rnn(x[:lengths[n], n:n + 1])       # zero initial state, both directions over the valid slice
                out[:lengths[n], n] = y[:, 0]
            x = out
        fc = self.fc.module
        mask = (torch.arange(T).unsqueeze(1) < torch.as_tensor(lengths).unsqueeze(0))
        rows = x[mask]
        if isinstance(fc, nn.Sequential):
            rows = fc[1](fc[0](rows))
        else:
            rows = fc(rows)
        logits = torch.zeros(T, N, self.num_class + 1, dtype=x.dtype)
        logits[mask] = rows
        if not self.training:
            return F.log_softmax(logits, dim=-1)
        return logits


def warp_ctc_loss(activations, targets_1d, input_sizes, target_sizes, blank=0):
    """What warp-ctc's CTCLoss() (size_average=False) returns for [T, N, C] activations: the summed negative log likelihood
    of softmax(activations); targets concatenated 1-D."""
    lp = F.log_softmax(activations, dim=-1)
    return F.ctc_loss(lp, targets_1d.long(), torch.as_tensor(input_sizes).long(), torch.as_tensor(target_sizes).long(),
                      blank=blank, reduction="sum", zero_infinity=False)


def synthetic_packed_batch(T, N, F_, num_class, S, seed):
    """Zero-padded [T, N, F] features with descending lengths (create_RNN_input sorts by length), 1-D targets."""
    g = torch.Generator().manual_seed(seed)
    lengths = sorted([int(v) for v in torch.randint(max(2 * S + 2, T // 2), T + 1, (N,), generator=g)], reverse=True)
    lengths[0] = T
    x = torch.randn(T, N, F_, generator=g)
    for n in range(N):
        x[lengths[n]:, n] = 0.0
    tsz = [int(v) for v in torch.randint(max(1, S // 2), S + 1, (N,), generator=g)]
    targets = torch.cat([torch.randint(1, num_class + 1, (t,), generator=g) for t in tsz]).int()
    return x, lengths, targets, tsz
