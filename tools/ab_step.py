"""A/B of one environment switch on the training step: python tools/ab_step.py ENV_NAME [cfg] -> medians with ENV=0/1 and the
relative L2 difference of the gradients."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from ctc_pytorch_b200.model import CTC_Model
from ctc_pytorch_b200.loss import CTCLoss
from ctc_pytorch_b200 import ops

env = sys.argv[1]
name = sys.argv[2] if len(sys.argv) > 2 else "cfg2"
vals = sys.argv[3].split(",") if len(sys.argv) > 3 else ["0", "1"]
dev = "cuda"
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


g = {}
for mode in (vals[0], vals[1], vals[0], vals[1]):
    os.environ[env] = mode
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
    g[mode] = torch.cat([p.grad.flatten().double() for p in m.parameters()])
    print("%s %s=%s: fwd+loss+bwd median %.3f ms (min %.3f), loss %.6f" % (name, env, mode, ts[len(ts) // 2], ts[0], float(loss.detach())),
          flush=True)
print("%s: gradient rel L2 (%s=1 vs 0) %.3e, finite %s" % (name, env, float((g[vals[0]] - g[vals[1]]).norm() / g[vals[0]].norm()),
                                                          bool(torch.isfinite(g[vals[1]]).all())))
