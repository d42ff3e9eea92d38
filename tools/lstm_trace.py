"""Per-phase cycle trace of the forward recurrence (CTCB200_LSTM_TRACE=1) at the cfg2 shape, pipelined and plain kernels."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["CTCB200_LSTM_TRACE"] = "1"
import torch

import bench
from ctc_pytorch_b200.model import CTC_Model

cfg = dict(bench.CFG[sys.argv[1] if len(sys.argv) > 1 else "cfg2"])
cfg["L"] = 2
torch.manual_seed(0)
m = CTC_Model(rnn_param=bench.rnn_param(cfg), num_class=cfg["C"], drop_out=0.0).to("cuda")
x = bench.make_batch(cfg, 1)[0].to("cuda")
m.eval()
for mode in ("1", "0"):
    os.environ["CTCB200_LSTM_PIPE"] = mode
    sys.stderr.write("--- CTCB200_LSTM_PIPE=%s\n" % mode)
    with torch.no_grad():
        m(x)
    torch.cuda.synchronize()
