import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.test_gpu_parity import _golden_model, relnorm
from ctc_pytorch_b200.loss import CTCLoss
from ctc_pytorch_b200 import ops
gd = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
m, meta, g = _golden_model(gd, "cnn_rnn")
x = torch.from_numpy(g["x"]).cuda(); m.train()
out = m(x)
il = torch.from_numpy(g["input_lengths"]).cuda()
loss = CTCLoss(reduction="sum")(out, torch.from_numpy(g["targets"]).cuda(), il, torch.from_numpy(g["target_lengths"]).cuda()) / out.shape[1]
loss.backward()
print("loss", loss.item(), float(g["loss"]), "out maxabs", (out.detach().cpu() - torch.from_numpy(g["out_train"])).abs().max().item())
for k, p in m.named_parameters():
    step = meta["grad_step"][k]
    vals = p.grad.detach().cpu().reshape(-1)[::step][:256]
    ref = torch.from_numpy(g["gradvals/" + k])
    print("%-36s rel=%.3e norm=%.3e refnorm=%.3e" % (k, relnorm(vals, ref), p.grad.norm().item(), meta["grad_norm"][k]))
torch.manual_seed(0)
M, N, K = 4096, 320, 25600
a = torch.randn(M, K).bfloat16().cuda(); b = torch.randn(N, K).bfloat16().cuda()
ref = a.float() @ b.float().t(); ref64 = (a.double() @ b.double().t())
for tile in (0, 64, 128, 256):
    c = ops.gemm_tn(a, b, tile_n=tile)
    print("gemm tile", tile, "vs fp32 torch", relnorm(c, ref), "vs fp64", relnorm(c, ref64), "torch fp32 vs fp64", relnorm(ref, ref64))
