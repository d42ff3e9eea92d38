"""One training step inside a cudaProfilerStart/Stop range (for ncu --profile-from-start off).
usage: python tools/profile_step.py [cfg2|cfg3|cfg4] [T] [bf16|x3]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn as nn
import bench
from ctc_pytorch_b200.model import CTC_Model
from ctc_pytorch_b200.loss import CTCLoss
from ctc_pytorch_b200 import ops, synth

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
cfg = dict(bench.CFG[name])
if len(sys.argv) > 2: cfg["T"] = int(sys.argv[2])
dev = "cuda"
torch.manual_seed(0)
m = CTC_Model(**synth.model_kwargs(cfg)).to(dev)
m.precision = sys.argv[3] if len(sys.argv) > 3 else "bf16"
opt = torch.optim.Adam(m.parameters(), lr=1e-3, weight_decay=0.005, fused=True)
x, frac, tg, tl = (t.to(dev) for t in bench.make_batch(cfg, 1))
lossf = CTCLoss(reduction="sum"); m.train()
def step():
    out = m(x); il = (frac * out.shape[0]).long()
    loss = lossf(out, tg, il, tl) / x.shape[0]
    ops.greedy_decode(out, il)
    opt.zero_grad(); loss.backward(); opt.step()
for _ in range(2): step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
