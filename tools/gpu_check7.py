"""Overlapped weight-gradient pipeline: step time and gradient agreement with overlap off / on (cfg2, cfg4)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from ctc_pytorch_b200.model import CTC_Model
from ctc_pytorch_b200.loss import CTCLoss
from ctc_pytorch_b200 import ops

dev = "cuda"
names = sys.argv[1:] or ["cfg2", "cfg4"]
for name in names:
    cfg = dict(bench.CFG[name])
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

    res = {}
    for mode in ("0", "1", "0", "1"):
        os.environ["CTCB200_OVERLAP_WGRAD"] = mode
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        ts = []
        for _ in range(8):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            loss = step()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        g = torch.cat([p.grad.flatten().double() for p in m.parameters()])
        res.setdefault(mode, []).append((ts[len(ts) // 2], ts[0], float(loss), g))
        print("%s overlap=%s: fwd+loss+bwd median %.3f ms  min %.3f ms  loss %.6f" % (name, mode, ts[len(ts) // 2], ts[0], float(loss)),
              flush=True)
    g0, g1 = res["0"][0][3], res["1"][0][3]
    rel = float((g0 - g1).norm() / g0.norm())
    print("%s gradient agreement overlap on vs off: rel L2 %.3e, finite %s" % (name, rel, bool(torch.isfinite(g1).all())), flush=True)
    assert rel < 1e-4
    # experiment: single-MUFU activations in the forward recurrence
    os.environ["CTCB200_OVERLAP_WGRAD"] = "1"
    os.environ["CTCB200_LSTM_ACT"] = "approx"
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
    ga = torch.cat([p.grad.flatten().double() for p in m.parameters()])
    print("%s act=approx: median %.3f ms min %.3f ms loss %.6f (exact %.6f) grad rel L2 vs exact %.3e" % (
        name, ts[len(ts) // 2], ts[0], float(loss), res["1"][0][2], float((ga - g1).norm() / g1.norm())), flush=True)
    del os.environ["CTCB200_LSTM_ACT"]
    del m, x, flush
    torch.cuda.empty_cache()
print("gpu_check7 done")
