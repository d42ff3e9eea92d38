"""Recurrent kernels of the working tree against the last committed build (tools/_ab/libctcb200_head.so) on the same box,
alternating, bare C-ABI calls timed with CUDA events: guards the per-step chain against regressions from host-visible features
(streamed input projection, exclusive shared memory, tagged waits).  python tools/rec_ab_head.py [T N H]"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
import torch

T, N, H = (int(v) for v in (sys.argv[1:4] if len(sys.argv) >= 4 else (800, 32, 512)))
dev = "cuda"
torch.manual_seed(0)
R = T * N
P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
S = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
w32 = 0.04 * torch.randn(8 * H, H, device=dev)
whh = w32.bfloat16()
gx = torch.randn(R, 8 * H, device=dev)
hout = torch.empty(R, 2 * H, device=dev)
c_save = torch.empty(R, 2 * H, device=dev)
gates16 = torch.empty(R, 2 * H, 4, dtype=torch.float16, device=dev)
dh = torch.randn(R, 2 * H, device=dev)
dg = torch.empty(R, 8 * H, dtype=torch.bfloat16, device=dev)
scratch = torch.empty(64 << 20, dtype=torch.uint8, device=dev)
libs = {"head": ctypes.CDLL(os.path.join(ROOT, "tools", "_ab", "libctcb200_head.so")),
        "tree": ctypes.CDLL(os.path.join(ROOT, "ctc_pytorch_b200", "libctcb200.so"))}
for f in sorted(os.listdir(os.path.join(ROOT, "tools", "_ab"))):    # optional experiment builds (-DAB_...)
    if f.startswith("libctcb200_AB_"):
        libs[f[len("libctcb200_"):-3]] = ctypes.CDLL(os.path.join(ROOT, "tools", "_ab", f))
for L in libs.values():
    L.ctcb200_last_error.restype = ctypes.c_char_p


def chk(rc, L):
    if rc != 0:
        raise RuntimeError(L.ctcb200_last_error().decode())


def fwd(L):
    chk(L.ctcb200_lstm_fwd(P(gx), P(whh), None, P(hout), P(c_save), P(gates16), P(scratch), T, N, H, 0, 0, S()), L)


def bwd(L):
    chk(L.ctcb200_lstm_bwd(P(dh), P(whh), None, P(c_save), P(gates16), P(dg), None, None, None, P(scratch), T, N, H, 0, 0, None, None,
                           None, None, S()), L)


def timed(fn, reps=8):
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


out = {"shape": {"T": T, "N": N, "H": H}}
for rnd in range(3):
    for nm, L in libs.items():
        out.setdefault("fwd_ms_" + nm, []).append(round(timed(lambda: fwd(L)), 4))
        out.setdefault("bwd_ms_" + nm, []).append(round(timed(lambda: bwd(L)), 4))
print(json.dumps(out))
