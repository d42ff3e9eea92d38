"""cfg3: add_cnn=True — 2 x Conv2d front (ch (1,32),(32,32), k 3x3, stride (1,2),(2,2), pad (1,1)) + 4 x BiLSTM-512,
T=800 -> T'=400, N=32: training-step time on the GPU next to the oracle's CPU composition on a bounded sample."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn

import bench
from ctc_pytorch_b200.model import CTC_Model
from ctc_pytorch_b200.loss import CTCLoss
from ctc_pytorch_b200 import ops
from oracle.model_ref import RefAcousticModel

cfg = dict(bench.CFG["cfg2"])
layers = [[(1, 32), (3, 3), (1, 2), (1, 1), None], [(32, 32), (3, 3), (2, 2), (1, 1), None]]
cnn_param = {"layer": layers, "batch_norm": True, "activate_function": nn.ReLU}
dev = "cuda"
torch.manual_seed(0)
m = CTC_Model(add_cnn=True, cnn_param=cnn_param, rnn_param=bench.rnn_param(cfg), num_class=cfg["C"], drop_out=0.0).to(dev)
opt = torch.optim.Adam(m.parameters(), lr=1e-3, weight_decay=0.005, fused=True)
x, frac, tg, tl = (t.to(dev) for t in bench.make_batch(cfg, 1))
lossf = CTCLoss(reduction="sum")
m.train()
flush = torch.empty(192 << 20, dtype=torch.uint8, device=dev)


def step():
    out = m(x)
    il = (frac * out.shape[0]).long()
    loss = lossf(out, tg, il, tl) / x.shape[0]
    ops.greedy_decode(out, il)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()
    return loss, out.shape


for _ in range(3):
    step()
torch.cuda.synchronize()
ts = []
for _ in range(10):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); loss, shape = step(); e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts.sort()
ms = ts[len(ts) // 2]

# per-entry-point device time (one stream, overlap off), as bench.py does for cfg2
from ctc_pytorch_b200 import _lib
L = _lib.lib()
per_call, orig_call = {}, L.call


def timed_call(name, *a):
    s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s_.record(); r = orig_call(name, *a); e_.record()
    per_call.setdefault(name, []).append((s_, e_))
    return r


L.call = timed_call
m.overlap_wgrad = False
for _ in range(2):
    step()
torch.cuda.synchronize()
L.call = orig_call
m.overlap_wgrad = True
kern = {k: sum(a.elapsed_time(b) for a, b in v) / 2 for k, v in per_call.items()}
kern_ms = {k: round(v, 4) for k, v in sorted(kern.items(), key=lambda kv: -kv[1])}
kern_n = {k: len(v) // 2 for k, v in per_call.items()}

# CPU: the oracle's restatement of the reference composition, 2 utterances, 16 threads
threads = bench.cpu_threads()
torch.set_num_threads(threads)
ref = RefAcousticModel(cfg["F"], cfg["H"], cfg["L"], cfg["C"], batch_norm=True, cnn_layers=layers)
ropt = torch.optim.Adam(ref.parameters(), lr=1e-3, weight_decay=0.005)
xs, fs, tgs, tls = (t[:2].cpu() for t in (x, frac, tg, tl))
ref.train()
t0 = None
for it in range(2):
    if it == 1:
        t0 = time.perf_counter()
    out = ref(xs)
    il = (fs * out.shape[0]).long()
    l = nn.CTCLoss(reduction="sum")(out, tgs, il, tls) / 2
    ropt.zero_grad(); l.backward(); ropt.step()
cpu_s = time.perf_counter() - t0
res = {"config": "cfg3: 2xConv2d front + 4xBiLSTM-512, T=800 -> T'=%d, N=32, C=62" % shape[0], "gpu_ms_per_step": ms,
       "gpu_utt_s": cfg["N"] / (ms * 1e-3), "loss": float(loss.detach()), "kernel_ms_per_step": kern_ms, "calls_per_step": kern_n,
       "cpu": {"kind": "port", "cores": threads, "utt_s": 2 / cpu_s, "sample": "2 utterances x 1 timed step"}}
print(json.dumps(res))
if len(sys.argv) > 1:
    json.dump(res, open(sys.argv[1], "w"), indent=1)
