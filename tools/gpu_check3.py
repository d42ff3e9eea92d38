import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ctc_pytorch_b200 import ops
dev = "cuda"
def t(M, N, K, lda, ldb, ak=0, bk=0):
    a = torch.randn(M, lda, device=dev).bfloat16(); b = torch.randn(N, ldb, device=dev).bfloat16()
    ref = a[:, ak:ak+K].float() @ b[:, bk:bk+K].float().t()
    try:
        c = ops.gemm_tn(a, b, a_koff=ak, b_koff=bk, k=K); torch.cuda.synchronize()
        print("M=%d N=%d K=%d lda=%d ldb=%d koff=(%d,%d) err=%.3e" % (M, N, K, lda, ldb, ak, bk, (c - ref).abs().max().item()), flush=True)
    except Exception as e:
        print("M=%d N=%d K=%d lda=%d ldb=%d koff=(%d,%d) EXC %s" % (M, N, K, lda, ldb, ak, bk, str(e).split("\n")[0]), flush=True)
        sys.exit(1)
t(18, 1024, 40, 40, 40)
t(18, 10, 256, 256, 256)
t(16, 256, 24, 24, 24)
t(10, 256, 24, 24, 24)
t(10, 256, 16, 24, 24)
t(10, 256, 18, 24, 24)
t(18, 256, 10, 16, 16)
t(1024, 40, 18, 24, 24)
t(512, 128, 15, 24, 24, 3, 0)
t(512, 128, 15, 24, 24, 0, 3)
