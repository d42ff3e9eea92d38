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


# ---------------------------------------------------------------------------------------------------------------------
# Design check for the CUDA path of N1: "alignment instead of masks". The recurrent kernels always scan all T
# rows of a padded batch; packed semantics follow if the forward direction sees the LEFT-aligned batch and the reverse
# direction a RIGHT-aligned copy. aligned_bilstm() emulates exactly what those kernels would compute (full-length scans,
# arbitrary garbage in the padding rows) so that tests can compare it with the per-utterance formulation above.
# ---------------------------------------------------------------------------------------------------------------------
def _scan(w_ih, w_hh, x, reverse):
    """Plain LSTM scan over all rows of x [T, N, I] (zero initial state, no bias), t = 0..T-1 or T-1..0; returns h [T, N, H]."""
    T, N, _ = x.shape
    H = w_hh.shape[1]
    h = x.new_zeros(N, H)
    c = x.new_zeros(N, H)
    out = []
    order = range(T - 1, -1, -1) if reverse else range(T)
    for t in order:
        g = x[t] @ w_ih.t() + h @ w_hh.t()
        i, f, gg, o = g.chunk(4, dim=1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * torch.tanh(c)
        out.append(h)
    if reverse:
        out.reverse()
    return torch.stack(out, 0)


def right_aligned(x, lengths, fill=None):
    """x [T, N, .] left-aligned -> frame k of utterance n at row T - len_n + k; padding rows = `fill` rows (or zeros)."""
    T = x.shape[0]
    out = torch.zeros_like(x) if fill is None else fill.clone()
    for n, L in enumerate(lengths):
        out[T - L:, n] = x[:L, n]
    return out


def left_aligned(x_r, lengths):
    """Inverse of right_aligned; padding rows are zero."""
    T = x_r.shape[0]
    out = torch.zeros_like(x_r)
    for n, L in enumerate(lengths):
        out[:L, n] = x_r[T - L:, n]
    return out


def aligned_bilstm(rnn, x, lengths, garbage=None):
    """Bidirectional layer output [T, N, 2H] with packed semantics, computed the way the CUDA path would: two full-length
    scans (forward over the left-aligned batch, reverse over the right-aligned one), realignment and zeroing of the padding.
    `garbage` [T, N, I]: what the padding rows contain (anything finite) — the result must not depend on it."""
    T, N, _ = x.shape
    lengths = [int(v) for v in lengths]
    mask = (torch.arange(T).unsqueeze(1) < torch.as_tensor(lengths).unsqueeze(0)).unsqueeze(-1)
    x_l = x if garbage is None else torch.where(mask, x, garbage)
    h_f = _scan(rnn.weight_ih_l0, rnn.weight_hh_l0, x_l, reverse=False)
    x_r = right_aligned(x, lengths, fill=garbage)
    h_r = left_aligned(_scan(rnn.weight_ih_l0_reverse, rnn.weight_hh_l0_reverse, x_r, reverse=True), lengths)
    return torch.cat([h_f * mask, h_r], dim=-1)


def _count_bn(bn, x, n_valid, valid_mask):
    """BatchNorm1d over rows whose padding is ZERO: the column sums over all T*N rows equal the sums over the valid rows, only
    the divisor is the valid-frame count (biased variance for normalisation, unbiased for running_var, as torch does).
    Output rows at padded positions are re-zeroed (BN(0) = shift is not zero)."""
    T, N, C = x.shape
    flat = x.reshape(T * N, C)
    if bn.training:
        s1, s2 = flat.sum(0), (flat * flat).sum(0)
        mean = s1 / n_valid
        var = s2 / n_valid - mean * mean
        if bn.track_running_stats:
            with torch.no_grad():
                m = bn.momentum
                bn.running_mean.mul_(1 - m).add_(m * mean.detach())
                bn.running_var.mul_(1 - m).add_(m * var.detach() * n_valid / max(n_valid - 1, 1))
                bn.num_batches_tracked.add_(1)
    else:
        mean, var = bn.running_mean, bn.running_var
    y = (flat - mean) * torch.rsqrt(var + bn.eps) * bn.weight + bn.bias
    return y.reshape(T, N, C) * valid_mask


def aligned_model_forward(model, x, lengths, garbage_scale=0.0):
    """RefPackedModel.forward restated the way the CUDA path of N1 is planned: dense [T, N, .] tensors with zero
    padding, BatchNorm by column sums + valid count, the recurrent layers by aligned_bilstm, the output layer on all rows with
    the padding re-zeroed. Must equal model(x, lengths) — tests/test_oracle.py."""
    T, N, _ = x.shape
    lengths = [int(v) for v in lengths]
    n_valid = sum(lengths)
    mask = (torch.arange(T).unsqueeze(1) < torch.as_tensor(lengths).unsqueeze(0)).unsqueeze(-1).to(x.dtype)
    for layer in model.rnns.children():
        if layer.batch_norm is not None:
            x = _count_bn(layer.batch_norm.module, x, n_valid, mask)
        garbage = garbage_scale * torch.randn_like(x) if garbage_scale else None
        x = aligned_bilstm(layer.rnn, x, lengths, garbage=garbage)
    fc = model.fc.module
    if isinstance(fc, nn.Sequential):
        x = _count_bn(fc[0], x, n_valid, mask)
        logits = (x.reshape(T * N, -1) @ fc[1].weight.t()).reshape(T, N, -1) * mask
    else:
        logits = (x.reshape(T * N, -1) @ fc.weight.t()).reshape(T, N, -1) * mask
    if not model.training:
        return F.log_softmax(logits, dim=-1)
    return logits
