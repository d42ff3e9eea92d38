"""Pipelined BPTT kernel (CTCB200_LSTM_PIPE_BWD=1, default) against the un-pipelined one: gradients, step time."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from ctc_pytorch_b200.model import CTC_Model
from ctc_pytorch_b200.loss import CTCLoss
from ctc_pytorch_b200 import ops

dev = "cuda"
lossf = CTCLoss(reduction="sum")


def make(name, T=None, N=None):
    cfg = dict(bench.CFG[name])
    if T: cfg["T"] = T
    if N: cfg["N"] = N
    torch.manual_seed(0)
    m = CTC_Model(rnn_param=bench.rnn_param(cfg), num_class=cfg["C"], drop_out=0.0).to(dev)
    x, frac, tg, tl = (t.to(dev) for t in bench.make_batch(cfg, 1))
    m.train()

    def step():
        out = m(x)
        il = (frac * out.shape[0]).long()
        loss = lossf(out, tg, il, tl) / x.shape[0]
        ops.greedy_decode(out, il)
        m.zero_grad(set_to_none=True)
        loss.backward()
        return loss
    return m, step


for name, T, N in (("cfg1", 200, 4), ("cfg2", 200, 21), ("cfg2", 800, 32)):
    m, step = make(name, T, N)
    g = {}
    for mode in ("0", "1"):
        os.environ["CTCB200_LSTM_PIPE_BWD"] = mode
        step()
        torch.cuda.synchronize()
        g[mode] = torch.cat([p.grad.flatten().double() for p in m.parameters()])
    rel = float((g["0"] - g["1"]).norm() / g["0"].norm())
    print("%s T=%d N=%d: grad rel L2 pipe vs plain %.3e finite %s" % (name, T, N, rel, bool(torch.isfinite(g["1"]).all())), flush=True)
    assert rel < 5e-3
    del m

m, step = make("cfg2")
flush = torch.empty(192 << 20, dtype=torch.uint8, device=dev)
for mode in ("0", "1", "0", "1"):
    os.environ["CTCB200_LSTM_PIPE_BWD"] = mode
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
    print("cfg2 bwd pipe=%s: fwd+loss+bwd median %.3f ms (min %.3f), loss %.6f" % (mode, ts[len(ts) // 2], ts[0], float(loss.detach())), flush=True)
print("gpu_check9 done")
