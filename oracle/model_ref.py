"""ORACLE — test infrastructure only (never imported by the product path).

CPU restatement of the reference's acoustic model, timit/models/model_ctc.py:13-185, on top of the same
third-party library the reference itself calls for every FLOP (PyTorch: nn.LSTM / nn.BatchNorm / nn.Linear /
nn.Conv2d / log_softmax; un-vendored, unpinned in requirements.txt, "pytorch1.2" per README.md:1-2; the
version here is the installed torch 2.11.0 running its CPU kernels). The module tree reproduces the
reference's state_dict keys (`conv.{n}.conv|batch_norm`, `rnns.{l}.batch_norm|rnn`, `fc.0|fc.1` or `fc`) so
one set of weights can be loaded into the reference, this oracle and the CUDA model alike.

Semantics restated:
  * input [N,T,F] -> time-major; LSTM(bias=False, bidirectional) over all padded frames, h0=c0=0
  * layers >= 1 (and the output layer) are preceded by BatchNorm1d whose statistics run over the T*N rows
  * layer 0 never has BatchNorm; dropout follows every LSTM layer
  * CNN front: Conv2d(bias) -> BatchNorm2d -> activation -> optional MaxPool2d -> dropout per block, then
    [N,Cc,T',F'] -> [T',N,Cc*F']
  * output: Linear(no bias) -> log_softmax over classes, shape [T',N,C]

Pinned by tests/test_oracle.py: agreement to float32 round-off (2e-6) with the reference's own CTC_Model (imported from
/root/reference/timit when present) under a shared state_dict, and by the committed golden vectors.
"""
from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F


class _RnnBlock(nn.Module):
    def __init__(self, input_size, hidden, batch_norm, dropout, bidirectional=True, rnn_type=nn.LSTM):
        super().__init__()
        self.batch_norm = nn.BatchNorm1d(input_size) if batch_norm else None
        self.rnn = rnn_type(input_size=input_size, hidden_size=hidden, bidirectional=bidirectional, bias=False)
        self.p = dropout
        self.fixed_mask = None   # tests: keep-mask [T*N, 2H] shared with the CUDA path instead of this process's RNG

    def forward(self, seq):  # seq [T, N, I]
        if self.batch_norm is not None:
            T, N, I = seq.shape
            seq = self.batch_norm(seq.reshape(T * N, I)).reshape(T, N, I)
        out, _ = self.rnn(seq)
        if self.fixed_mask is not None and self.training and self.p > 0:
            return out * self.fixed_mask.view_as(out).to(out.dtype) / (1.0 - self.p)
        return F.dropout(out, self.p, self.training)


class _ConvBlock(nn.Module):
    def __init__(self, cin, cout, kernel, stride, padding, pool, batch_norm, act, dropout):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, kernel_size=kernel, stride=stride, padding=padding)
        self.batch_norm = nn.BatchNorm2d(cout) if batch_norm else None
        self.act = act()
        self.pool = pool
        self.p = dropout

    def forward(self, x):
        x = self.conv(x)
        if self.batch_norm is not None:
            x = self.batch_norm(x)
        x = self.act(x)
        if self.pool is not None:
            x = F.max_pool2d(x, self.pool)
        return F.dropout(x, self.p, self.training)


class RefAcousticModel(nn.Module):
    def __init__(self, input_size, hidden, layers, num_class, batch_norm=True, cnn_layers=None, cnn_batch_norm=True,
                 cnn_act=nn.ReLU, dropout=0.0, bidirectional=True, rnn_type=nn.LSTM):
        super().__init__()
        D = 2 if bidirectional else 1
        rnn_in = input_size
        self.has_cnn = bool(cnn_layers)
        if self.has_cnn:
            blocks = []
            cout = 1
            for n, ((cin, cout), kernel, stride, padding, pool) in enumerate(cnn_layers):
                blocks.append((str(n), _ConvBlock(cin, cout, kernel, stride, padding, pool, cnn_batch_norm, cnn_act, dropout)))
                rnn_in = (rnn_in + 2 * padding[1] - kernel[1]) // stride[1] + 1
            self.conv = nn.Sequential(OrderedDict(blocks))
            rnn_in *= cout
        blocks = [("0", _RnnBlock(rnn_in, hidden, False, dropout, bidirectional, rnn_type))]
        for l in range(1, layers):
            blocks.append((str(l), _RnnBlock(D * hidden, hidden, batch_norm, dropout, bidirectional, rnn_type)))
        self.rnns = nn.Sequential(OrderedDict(blocks))
        if batch_norm:
            self.fc = nn.Sequential(nn.BatchNorm1d(D * hidden), nn.Linear(D * hidden, num_class, bias=False))
        else:
            self.fc = nn.Linear(D * hidden, num_class, bias=False)

    def forward(self, x):  # x [N, T, F]
        if self.has_cnn:
            y = self.conv(x.unsqueeze(1))                      # [N, Cc, T', F']
            N, Cc, Tp, Fp = y.shape
            seq = y.permute(2, 0, 1, 3).reshape(Tp, N, Cc * Fp)
        else:
            seq = x.transpose(0, 1)
        seq = self.rnns(seq)
        T, N, D = seq.shape
        logits = self.fc(seq.reshape(T * N, D)).reshape(T, N, -1)
        return F.log_softmax(logits, dim=-1)


from ctc_pytorch_b200.synth import synthetic_batch  # noqa: E402,F401  (one generator for bench, tests and golden scripts)
