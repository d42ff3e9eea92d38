"""Per-phase cycle trace of the pipelined BPTT kernel (CTCB200_LSTM_TRACE=1) at the cfg2 shape."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ctc_pytorch_b200 import _lib

T, N, H = 800, 32, 512
dev = "cuda"
torch.manual_seed(0)
L = _lib.lib()
R = T * N
whhT = (0.05 * torch.randn(8 * H, H, device=dev)).to(torch.bfloat16)
c_save = torch.randn(R, 2 * H, device=dev)
gates = torch.rand(R, 2 * H, 4, device=dev).to(torch.float16)
dh = torch.randn(R, 2 * H, device=dev)
scratch = torch.empty(L.dll.ctcb200_lstm_scratch_bytes(N, H), dtype=torch.uint8, device=dev)
dg = torch.empty(R, 8 * H, dtype=torch.bfloat16, device=dev)
for mode in ("1", "0"):
    os.environ["CTCB200_LSTM_PIPE_BWD"] = mode
    os.environ.pop("CTCB200_LSTM_TRACE", None)
    for _ in range(2):
        L.call("ctcb200_lstm_bwd", _lib.ptr(dh), _lib.ptr(whhT), _lib.ptr(c_save), _lib.ptr(gates), _lib.ptr(dg), _lib.ptr(scratch),
               T, N, H, 0, None, None, None, None, _lib.stream())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    L.call("ctcb200_lstm_bwd", _lib.ptr(dh), _lib.ptr(whhT), _lib.ptr(c_save), _lib.ptr(gates), _lib.ptr(dg), _lib.ptr(scratch),
           T, N, H, 0, None, None, None, None, _lib.stream())
    e1.record()
    torch.cuda.synchronize()
    print("pipe_bwd=%s: %.3f ms per launch (%.3f us/step)" % (mode, e0.elapsed_time(e1), e0.elapsed_time(e1) * 1e3 / T), flush=True)
    if mode == "1":
        os.environ["CTCB200_LSTM_TRACE"] = "1"
        L.call("ctcb200_lstm_bwd", _lib.ptr(dh), _lib.ptr(whhT), _lib.ptr(c_save), _lib.ptr(gates), _lib.ptr(dg), _lib.ptr(scratch),
               T, N, H, 0, None, None, None, None, _lib.stream())
        torch.cuda.synchronize()
