"""Small end-to-end run of every kernel family for `compute-sanitizer` (memcheck / racecheck / synccheck): a tiny padded model
step in both precision modes (cluster recurrent kernels with DSMEM bulk copies, GEMMs, BatchNorm, log-softmax), the CNN front,
the packed model, CTC loss, greedy / beam decode, edit distance. Shapes are CI-sized so the instrumented run finishes in minutes.

    compute-sanitizer --tool memcheck --log-file gpurun_out/memcheck_r2.log python tools/sanitize_target.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.nn as nn

from ctc_pytorch_b200 import ops, synth
from ctc_pytorch_b200.decoder import BeamDecoder, GreedyDecoder
from ctc_pytorch_b200.loss import CTCLoss
from ctc_pytorch_b200.model import CTC_Model
from ctc_pytorch_b200.packed import CTC_RNN, WarpCTCLoss

dev = "cuda"
which = sys.argv[1] if len(sys.argv) > 1 else "all"
torch.manual_seed(0)
T, N, F, H, L, C = 10, 5, 40, 128, 2, 12


def padded(precision, cnn):
    cfg = dict(T=T, N=N, F=F, C=C, H=H, L=L, S=4, cnn=cnn)
    m = CTC_Model(**synth.model_kwargs(cfg, drop_out=0.1)).to(dev)
    m.precision = precision
    x, frac, tg, tl = synth.synthetic_batch(T, N, F, C, 4, 1)
    m.train()
    out = m(x.to(dev))
    il = (frac.to(dev) * out.shape[0]).long()
    loss = CTCLoss(reduction="sum")(out, tg.to(dev), il, tl.to(dev)) / N
    loss.backward()
    torch.cuda.synchronize()
    return out.detach()


if which in ("all", "model"):
    for prec in ("bf16", "x3"):
        for cnn in (False, True):
            out = padded(prec, cnn)
            print("padded model", prec, "cnn" if cnn else "rnn", tuple(out.shape))
if which in ("all", "packed"):
    m = CTC_RNN(rnn_input_size=F, rnn_hidden_size=H, rnn_layers=2, num_class=C, drop_out=0.0).to(dev)
    lens = [T, T - 2, T - 3, 6, 5]
    x = torch.randn(T, N, F)
    for n, l in enumerate(lens):
        x[l:, n] = 0
    act = m(nn.utils.rnn.pack_padded_sequence(x.to(dev), lens))
    tg = torch.randint(1, C + 1, (10,), dtype=torch.int32)
    WarpCTCLoss()(act, tg.to(dev), lens, [2] * N).backward()
    torch.cuda.synchronize()
    print("packed model", tuple(act.shape))
if which in ("all", "decode"):
    units = ["blank", "UNK", "a", "b", "c", "d", "e", "f"]
    lp = torch.log_softmax(2 * torch.randn(30, 3, 8), -1).to(dev)
    print(GreedyDecoder(dict(enumerate(units)), space_idx=-1).decode(lp, [30, 25, 12]))
    dec = BeamDecoder(dict(enumerate(units)), beam_width=10, lm_path=os.path.join(ROOT, "tests", "golden", "lm_c8.arpa"), lm_alpha=0.1)
    print(dec.decode(lp, [30, 25, 12]))
    hyp = torch.randint(1, 6, (4, 20), dtype=torch.int32, device=dev)
    print(ops.edit_distance(hyp, torch.tensor([20, 10, 0, 5], dtype=torch.int32, device=dev),
                            torch.randint(1, 6, (4, 15), device=dev), torch.tensor([15, 7, 3, 0], device=dev)).tolist())
print("sanitize target done")
