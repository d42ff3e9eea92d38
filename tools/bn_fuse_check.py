"""BatchNorm backward: unfused kernel vs coefficient form vs float64, and BPTT with the coefficients applied on the fly vs
BPTT on the pre-applied gradient."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ctc_pytorch_b200 import _lib

dev = "cuda"
L = _lib.lib()
torch.manual_seed(0)
T, N, H = 64, 32, 512
R, C = T * N, 2 * H
x = torch.tanh(torch.randn(R, C, device=dev) * 0.7 + 0.2 * torch.randn(C, device=dev))
dy = 1e-4 * torch.randn(R, C, device=dev) * (1 + torch.randn(C, device=dev))
gamma = torch.rand(C, device=dev) + 0.5
x64, dy64 = x.double(), dy.double()
mean64, var64 = x64.mean(0), x64.var(0, unbiased=False)
rs64 = (var64 + 1e-5).rsqrt()
xh = (x64 - mean64) * rs64
s1, s2 = dy64.sum(0), (dy64 * xh).sum(0)
dx64 = gamma.double() * rs64 * (dy64 - s1 / R - xh * s2 / R)
mean, rstd = mean64.float(), rs64.float()
ws = torch.empty(2 * C, dtype=torch.float64, device=dev)
dx_un = torch.empty_like(dy)
dg1, db1 = torch.empty(C, device=dev), torch.empty(C, device=dev)
L.call("ctcb200_bn_bwd", _lib.ptr(dy), _lib.ptr(x), _lib.ptr(mean), _lib.ptr(rstd), _lib.ptr(gamma), _lib.ptr(dx_un), _lib.ptr(dg1),
       _lib.ptr(db1), R, C, _lib.ptr(ws), _lib.stream())
coef = torch.empty(3 * C, device=dev)
dg2, db2 = torch.empty(C, device=dev), torch.empty(C, device=dev)
L.call("ctcb200_bn_bwd_coef", _lib.ptr(dy), _lib.ptr(x), _lib.ptr(mean), _lib.ptr(rstd), _lib.ptr(gamma), _lib.ptr(coef), _lib.ptr(dg2),
       _lib.ptr(db2), R, C, _lib.ptr(ws), _lib.stream())
A, B, D = coef[:C], coef[C:2 * C], coef[2 * C:]
dx_fu = A * dy + B * x + D
rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
print("dx: unfused vs f64 %.2e | coefficient form vs f64 %.2e | dgamma %.2e %.2e | dbeta %.2e %.2e" % (
    rel(dx_un, dx64), rel(dx_fu, dx64), rel(dg1, s2), rel(dg2, s2), rel(db1, s1), rel(db2, s1)))

whhT = (0.05 * torch.randn(8 * H, H, device=dev)).to(torch.bfloat16)
c_save = torch.randn(R, C, device=dev)
gates = torch.rand(R, C, 4, device=dev).to(torch.float16)
scratch = torch.empty(L.dll.ctcb200_lstm_scratch_bytes(N, H), dtype=torch.uint8, device=dev)
out = []
for fused in (False, True):
    dg = torch.zeros(R, 8 * H, dtype=torch.bfloat16, device=dev)
    L.call("ctcb200_lstm_bwd", _lib.ptr(dy if fused else dx_un), _lib.ptr(whhT), _lib.ptr(c_save), _lib.ptr(gates), _lib.ptr(dg),
           _lib.ptr(scratch), T, N, H, 0, _lib.ptr(x) if fused else None, _lib.ptr(coef) if fused else None, None, None, _lib.stream())
    torch.cuda.synchronize()
    out.append(dg.float())
print("BPTT dG: fused vs pre-applied rel L2 %.2e, max abs %.2e (max |dG| %.2e)" % (rel(out[1], out[0]), float((out[1] - out[0]).abs().max()),
                                                                             float(out[0].abs().max())))
