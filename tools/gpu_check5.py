import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn as nn
import bench
from ctc_pytorch_b200.model import CTC_Model
cfg = dict(bench.CFG["cfg2"]); cfg["T"] = 300
torch.manual_seed(0)
m = CTC_Model(rnn_param=bench.rnn_param(cfg), num_class=cfg["C"], drop_out=0.0).cuda()
m.batch_tile = int(sys.argv[1]) if len(sys.argv) > 1 else 0
x = bench.make_batch(cfg, 1)[0].cuda()
m.train()
for _ in range(2):
    out = m(x); torch.cuda.synchronize()
