"""CTC loss kernels alone at a (T, N, C, S) shape: device time of the alpha/beta sweeps and of the gradient kernel."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ctc_pytorch_b200.loss import CTCLoss

T, N, C, S = (int(v) for v in (sys.argv[1:5] if len(sys.argv) >= 5 else (800, 32, 62, 60)))
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 10
dev = "cuda"
torch.manual_seed(0)
lp = torch.randn(T, N, C, device=dev).log_softmax(-1).requires_grad_(True)
tl = torch.randint(max(1, S // 2), S + 1, (N,), device=dev)
tg = torch.randint(1, C, (N, S), device=dev)
il = torch.randint(int(0.8 * T), T + 1, (N,), device=dev)
lossf = CTCLoss(reduction="sum")
for _ in range(3):
    lossf(lp, tg, il, tl).backward()
torch.cuda.synchronize()
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
tf = tb = 0.0
for _ in range(reps):
    e0.record(); l = lossf(lp, tg, il, tl); e1.record(); l.backward(); e2.record()
    torch.cuda.synchronize()
    tf += e0.elapsed_time(e1) / reps; tb += e1.elapsed_time(e2) / reps
ref = torch.nn.functional.ctc_loss(lp.detach().double().cpu(), tg.cpu(), il.cpu(), tl.cpu(), reduction="sum")
alg = 2.0 * T * N * C * 4
print("ctc T=%d N=%d C=%d S=%d: fwd (alpha+beta sweeps, incl. host wrapper) %.3f ms, bwd (grad) %.3f ms, loss rel err vs torch f64 %.2e, "
      "algorithmic %.1f MB -> %.1f GB/s over fwd+bwd" % (T, N, C, S, tf, tb, abs(float(l) - float(ref)) / abs(float(ref)), alg / 1e6,
                                                      alg / ((tf + tb) * 1e-3) / 1e9))
