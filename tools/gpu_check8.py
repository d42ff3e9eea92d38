"""Pipelined forward recurrence (CTCB200_LSTM_PIPE=1, default) against the un-pipelined kernel: identical outputs, step time."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from ctc_pytorch_b200.model import CTC_Model
from ctc_pytorch_b200.loss import CTCLoss
from ctc_pytorch_b200 import ops

dev = "cuda"
for name, T in (("cfg1", 64), ("cfg2", 50), ("cfg2", 800)):
    cfg = dict(bench.CFG[name]); cfg["T"] = T
    torch.manual_seed(0)
    m = CTC_Model(rnn_param=bench.rnn_param(cfg), num_class=cfg["C"], drop_out=0.0).to(dev)
    x, frac, tg, tl = (t.to(dev) for t in bench.make_batch(cfg, 1))
    m.eval()
    outs = {}
    for mode in ("0", "1"):
        os.environ["CTCB200_LSTM_PIPE"] = mode
        with torch.no_grad():
            outs[mode] = m(x).clone()
        torch.cuda.synchronize()
    d = (outs["0"] - outs["1"]).abs().max().item()
    print("%s T=%d: max |out(pipe) - out(plain)| = %.3e, finite %s" % (name, T, d, bool(torch.isfinite(outs["1"]).all())), flush=True)
    assert d < 1e-5

cfg = dict(bench.CFG["cfg2"])
torch.manual_seed(0)
m = CTC_Model(rnn_param=bench.rnn_param(cfg), num_class=cfg["C"], drop_out=0.0).to(dev)
x, frac, tg, tl = (t.to(dev) for t in bench.make_batch(cfg, 1))
lossf = CTCLoss(reduction="sum")
m.train()
flush = torch.empty(192 << 20, dtype=torch.uint8, device=dev)


def step():
    out = m(x)
    il = (frac * out.shape[0]).long()
    loss = lossf(out, tg, il, tl) / x.shape[0]
    ops.greedy_decode(out, il)
    m.zero_grad(set_to_none=True)
    loss.backward()
    return loss


for mode in ("0", "1", "0", "1"):
    os.environ["CTCB200_LSTM_PIPE"] = mode
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    ts = []
    for _ in range(8):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); loss = step(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    # forward-only time
    tf = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.no_grad():
            e0.record(); m(x); e1.record()
        torch.cuda.synchronize()
        tf.append(e0.elapsed_time(e1))
    print("cfg2 pipe=%s: fwd+loss+bwd median %.3f ms (min %.3f), inference fwd %.3f ms, loss %.6f" % (
        mode, ts[len(ts) // 2], ts[0], min(tf), float(loss.detach())), flush=True)
print("gpu_check8 done")
