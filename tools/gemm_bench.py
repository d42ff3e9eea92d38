"""GEMM shapes of the cfg2 / cfg4 training step: time (CUDA events, L2 flushed) and TFLOP/s, ours vs cuBLAS."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ctc_pytorch_b200 import ops
dev = "cuda"
flush = torch.empty(192 << 20, dtype=torch.uint8, device=dev)
def bench(fn, iters=10):
    fn(); torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
    return tot / iters
shapes = [("Gx l>=1   cfg2", 25600, 4096, 1024), ("Gx l0     cfg2", 25600, 4096, 40), ("dX        cfg2", 25600, 1024, 4096),
          ("dWih      cfg2", 4096, 1024, 25600), ("dWhh      cfg2", 2048, 512, 25568), ("fc        cfg2", 25600, 62, 1024),
          ("Gx l>=1   cfg4", 76800, 5120, 1280)]
for name, M, N, K in shapes:
    a = torch.randn(M, K + (8 - K % 8) % 8, device=dev).bfloat16(); b = torch.randn(N, a.shape[1], device=dev).bfloat16()
    ref = a[:, :K].float() @ b[:, :K].float().t()
    for tile in (0, 128, 256):
        if tile > 64 and N <= 64: continue
        c = ops.gemm_tn(a, b, k=K, tile_n=tile)
        err = ((c - ref).norm() / ref.norm()).item()
        ms = bench(lambda: ops.gemm_tn(a, b, out=c, k=K, tile_n=tile))
        print("%s M=%d N=%d K=%d tile=%3d: %.3f ms %7.1f TFLOP/s  relerr %.1e" % (name, M, N, K, tile, ms, 2.0 * M * N * K / ms / 1e9, err), flush=True)
    ab, bb = a[:, :K].contiguous(), b[:, :K].contiguous()
    ms = bench(lambda: torch.matmul(ab, bb.t()))
    print("%s cuBLAS bf16->bf16: %.3f ms %7.1f TFLOP/s" % (name, ms, 2.0 * M * N * K / ms / 1e9), flush=True)
