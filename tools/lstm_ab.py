"""A/B timing of the recurrent kernels at one layer's shape: the current library (bf16 and x3 operand modes) and, when
tools/_ab/libctcb200_r1.so exists, the round-1 library — CUDA events around the bare C-ABI calls.
usage: python tools/lstm_ab.py [T=800] [N=32] [H=512]"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

T, N, H = (int(v) for v in (sys.argv[1:4] if len(sys.argv) >= 4 else (800, 32, 512)))
dev = "cuda"
torch.manual_seed(0)
R = T * N
P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
S = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
w32 = 0.04 * torch.randn(8 * H, H, device=dev)
whh, whh_lo = w32.bfloat16(), (w32 - w32.bfloat16().float()).bfloat16()
gx = torch.randn(R, 8 * H, device=dev)
hout = torch.empty(R, 2 * H, device=dev)
c_save = torch.empty(R, 2 * H, device=dev)
gates16 = torch.empty(R, 2 * H, 4, dtype=torch.float16, device=dev)
gates32 = torch.empty(R, 2 * H, 4, device=dev)
dh = torch.randn(R, 2 * H, device=dev)
dg = torch.empty(R, 8 * H, dtype=torch.bfloat16, device=dev)
dg_lo = torch.empty(R, 8 * H, dtype=torch.bfloat16, device=dev)
dg2 = torch.empty(R, 8 * H, dtype=torch.bfloat16, device=dev)
dg2_lo = torch.empty(R, 8 * H, dtype=torch.bfloat16, device=dev)
scratch = torch.empty(64 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


res = {"shape": {"T": T, "N": N, "H": H}}
cur = ctypes.CDLL(os.path.join(ROOT, "ctc_pytorch_b200", "libctcb200.so"))
cur.ctcb200_last_error.restype = ctypes.c_char_p


def chk(rc, lib):
    if rc != 0:
        raise RuntimeError(lib.ctcb200_last_error().decode())


for name, lo, gates in (("bf16", None, gates16), ("x3", whh_lo, gates32)):
    for cell, cname in ((0, ""), (1, "_gru"), (2, "_rnn")):
        f = timed(lambda: chk(cur.ctcb200_lstm_fwd(P(gx), P(whh), P(lo), P(hout), P(c_save), P(gates), P(scratch), T, N, H, 0, cell,
                                                   S()), cur))
        b = timed(lambda: chk(cur.ctcb200_lstm_bwd(P(dh), P(whh), P(lo), P(c_save), P(gates), P(dg), P(dg_lo) if lo is not None else None,
                                                   P(dg2), P(dg2_lo) if lo is not None else None, P(scratch), T, N, H, 0, cell,
                                                   None, None, None, None, S()), cur))
        res["r2_" + name + cname] = {"fwd_ms": f, "bwd_ms": b, "fwd_us_per_step": f * 1e3 / T, "bwd_us_per_step": b * 1e3 / T}
os.environ["CTCB200_LSTM_PIPE"] = "0"
f = timed(lambda: chk(cur.ctcb200_lstm_fwd(P(gx), P(whh), None, P(hout), P(c_save), P(gates16), P(scratch), T, N, H, 0, 0, S()), cur))
res["r2_bf16_unpipelined_fwd_ms"] = f
os.environ.pop("CTCB200_LSTM_PIPE")
old_path = os.path.join(ROOT, "tools", "_ab", "libctcb200_r1.so")
if os.path.exists(old_path):
    old = ctypes.CDLL(old_path)
    old.ctcb200_last_error.restype = ctypes.c_char_p
    chk(cur.ctcb200_lstm_fwd(P(gx), P(whh), None, P(hout), P(c_save), P(gates16), P(scratch), T, N, H, 0, 0, S()), cur)
    f = timed(lambda: chk(old.ctcb200_lstm_fwd(P(gx), P(whh), P(hout), P(c_save), P(gates16), P(scratch), T, N, H, 0, S()), old))
    b = timed(lambda: chk(old.ctcb200_lstm_bwd(P(dh), P(whh), P(c_save), P(gates16), P(dg), P(scratch), T, N, H, 0, None, None, None, None,
                                               S()), old))
    res["r1_bf16"] = {"fwd_ms": f, "bwd_ms": b}
print(json.dumps(res))
