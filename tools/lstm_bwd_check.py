"""Kernel-level A/B of ctcb200_lstm_bwd: pipelined vs plain kernel on the same random inputs; prints where they differ."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ctc_pytorch_b200 import _lib

T, N, H = (int(v) for v in (sys.argv[1:4] if len(sys.argv) >= 4 else (12, 4, 256)))
dev = "cuda"
torch.manual_seed(0)
L = _lib.lib()
R = T * N
whhT = (0.05 * torch.randn(8 * H, H, device=dev)).to(torch.bfloat16)
c_save = torch.randn(R, 2 * H, device=dev)
gates = torch.rand(R, 2 * H, 4, device=dev).to(torch.float16)
dh = torch.randn(R, 2 * H, device=dev)
scratch = torch.empty(L.dll.ctcb200_lstm_scratch_bytes(N, H), dtype=torch.uint8, device=dev)
out = {}
res = torch.zeros(2, dtype=torch.int32, device=dev)
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 1
for mode in ("0", "1"):
    os.environ["CTCB200_LSTM_PIPE_BWD"] = mode
    for rep in range(reps):
        dg = torch.full((R, 8 * H), float("nan"), dtype=torch.bfloat16, device=dev)
        L.call("ctcb200_lstm_bwd", _lib.ptr(dh), _lib.ptr(whhT), _lib.ptr(c_save), _lib.ptr(gates), _lib.ptr(dg), _lib.ptr(scratch),
               T, N, H, 0, None, None, _lib.ptr(res), None, _lib.stream())
        torch.cuda.synchronize()
        if rep and not torch.equal(dg.float().nan_to_num(7.0), prev.float().nan_to_num(7.0)):
            print("mode %s: run %d differs from run 0 (non-deterministic)" % (mode, rep))
        if rep == 0:
            prev = dg
    out[mode] = prev.float().view(T, N, 2, H // 32, 32, 4)   # (t, n, dir, unit block, unit, gate)
a, b = out["0"], out["1"]
bad = ~torch.isfinite(b) | ((a - b).abs() > 1e-2 * (a.abs() + 1e-3))
print("T=%d N=%d H=%d: mismatching elements %d of %d; non-finite %d" % (T, N, H, int(bad.sum()), bad.numel(), int((~torch.isfinite(b)).sum())))
if bad.any():
    for name, dim in (("t", 0), ("n", 1), ("dir", 2), ("unit block", 3), ("unit", 4), ("gate", 5)):
        dims = [d for d in range(6) if d != dim]
        print("  by %-10s" % name, bad.sum(dim=dims).tolist())
    # BPTT order: dir 0 processes t = T-1 .. 0, dir 1 processes t = 0 .. T-1
    idx = bad.nonzero()[:6].tolist()
    for i in idx:
        print("  e.g.", i, float(a[tuple(i)]), float(b[tuple(i)]))
