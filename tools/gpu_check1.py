"""Dev harness (GPU): first-contact checks for CTC / greedy / tcgen05 GEMM with verbose diagnostics."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ctc_pytorch_b200 import ops, loss as L

torch.manual_seed(0)
dev = "cuda"

def check_gemm(M, N, K, tile_n=0, out_dtype=torch.float32, structured=False, a_koff=0, b_koff=0):
    if structured:
        a = torch.zeros(M, K + a_koff); b = torch.zeros(N, K + b_koff)
        for i in range(M): a[i, a_koff + (i % K)] = 1.0 + (i % 7)
        for j in range(N): b[j, b_koff:] = torch.arange(K, dtype=torch.float32) % 13 + j % 5
    else:
        a = torch.randn(M, K + a_koff); b = torch.randn(N, K + b_koff)
    a16 = a.to(torch.bfloat16).to(dev); b16 = b.to(torch.bfloat16).to(dev)
    ref = a16[:, a_koff:a_koff+K].float() @ b16[:, b_koff:b_koff+K].float().t()
    try:
        c = ops.gemm_tn(a16, b16, out_dtype=out_dtype, tile_n=tile_n, a_koff=a_koff, b_koff=b_koff, k=K)
        torch.cuda.synchronize()
    except Exception as e:
        print("GEMM M=%d N=%d K=%d tile=%d EXC %s" % (M, N, K, tile_n, e)); return False
    err = (c.float() - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-9
    ok = err <= (2e-2 if out_dtype == torch.bfloat16 else 1e-3) * scale
    print("GEMM M=%d N=%d K=%d tile=%d out=%s koff=(%d,%d) struct=%d maxerr=%.3e scale=%.3e %s" % (
        M, N, K, tile_n, str(out_dtype)[6:], a_koff, b_koff, structured, err, scale, "OK" if ok else "FAIL"))
    if not ok:
        d = (c.float() - ref).abs()
        bad = (d > 1e-2 * scale).nonzero()
        print("  bad count", bad.shape[0], "first", bad[:8].tolist())
        print("  c[0,:8]", c[0, :8].tolist()); print("  r[0,:8]", ref[0, :8].tolist())
        print("  c[1,:8]", c[1, :8].tolist()); print("  r[1,:8]", ref[1, :8].tolist())
        if M > 40: print("  c[40,:8]", c[40, :8].tolist(), " r", ref[40, :8].tolist())
    return ok

def check_ctc(T, N, C, S, seed=0, infeasible=False):
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(T, N, C, generator=g) * 2
    lp = torch.log_softmax(logits, -1)
    tl = torch.randint(max(1, S // 2), S + 1, (N,), generator=g)
    tl[0] = S
    if N > 2: tl[2] = 0
    tg = torch.zeros(N, S, dtype=torch.long)
    for i in range(N):
        tg[i, :tl[i]] = torch.randint(1, C, (int(tl[i]),), generator=g)
    if N > 1 and S > 3: tg[1, 1] = tg[1, 0]; tg[1, 2] = tg[1, 0]
    il = torch.linspace(1.0, 0.6, N).mul(T).round().long()
    if infeasible and N > 3: il[3] = 2
    lp_ref = lp.clone().requires_grad_(True)
    nll_ref = torch.nn.functional.ctc_loss(lp_ref, tg, il, tl, blank=0, reduction="none")
    mask = torch.isfinite(nll_ref)
    nll_ref[mask].sum().backward()
    lp_g = lp.to(dev).requires_grad_(True)
    nll = L.ctc_loss(lp_g, tg.to(dev), il.to(dev), tl.to(dev), blank=0, reduction="none")
    w = torch.zeros(N, device=dev); w[mask.to(dev)] = 1.0
    (nll * w)[mask.to(dev)].sum().backward()
    torch.cuda.synchronize()
    nl = nll.cpu()
    rel = ((nl[mask] - nll_ref[mask].detach()).abs() / nll_ref[mask].detach().abs().clamp(min=1e-6)).max().item()
    infok = bool((torch.isinf(nl) == torch.isinf(nll_ref.detach())).all())
    gg = lp_g.grad.cpu(); gr = lp_ref.grad
    gm = mask.view(1, N, 1).expand_as(gg)
    gerr = (gg[gm] - gr[gm]).abs().max().item()
    print("CTC T=%d N=%d C=%d S=%d nll_rel=%.2e inf_match=%s grad_abs=%.2e (gmax %.2e) %s" % (
        T, N, C, S, rel, infok, gerr, gr.abs().max().item(), "OK" if rel < 1e-4 and gerr < 1e-4 and infok else "FAIL"))

def check_greedy(T, N, C):
    lp = torch.log_softmax(torch.randn(T, N, C) * 3, -1)
    lp[5, 0, 1] = lp[5, 0, 4] = 1.0  # tie -> first index
    il = torch.linspace(1.0, 0.5, N).mul(T).long()
    idx, labels, ol = ops.greedy_decode(lp.to(dev), il.to(dev), blank=0)
    ref = lp.argmax(-1).t()
    ok = bool((idx.cpu().long() == ref).all())
    ok2 = True
    for i in range(N):
        seq = ref[i, :il[i]].tolist(); out = []
        for j, v in enumerate(seq):
            if v != 0 and (j == 0 or v != seq[j - 1]): out.append(v)
        got = labels[i, :ol[i]].cpu().tolist()
        ok2 = ok2 and got == out
    print("GREEDY T=%d N=%d C=%d argmax %s collapse %s" % (T, N, C, ok, ok2))

if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    import traceback
    for fn, args in [(check_greedy, (800, 32, 62)), (check_greedy, (50, 3, 5)), (check_ctc, (800, 32, 62, 60)),
                     (check_ctc, (100, 5, 10, 12, 1, True)), (check_ctc, (64, 4, 40, 100, 2)), (check_ctc, (30, 2, 6, 1, 3))]:
        try: fn(*args)
        except Exception: traceback.print_exc()
    allok = True
    for st in (True, False):
        allok &= check_gemm(128, 64, 64, 64, structured=st)
        allok &= check_gemm(128, 128, 64, 128, structured=st)
        allok &= check_gemm(128, 256, 128, 256, structured=st)
    for (M, N, K) in [(256, 256, 256), (1000, 300, 200), (25600, 4096, 1024), (25600, 4096, 40), (4096, 1024, 25600), (62, 1024, 2560), (2560, 62, 1024)]:
        for tn in (0, 64, 128, 256):
            if M * N * K > 5e10 and tn not in (0, 256): continue
            allok &= check_gemm(M, N, K, tn)
    allok &= check_gemm(512, 256, 992, 0, a_koff=32, b_koff=0)
    allok &= check_gemm(512, 256, 1000, 0, a_koff=8, b_koff=16)
    allok &= check_gemm(300, 200, 128, 0, out_dtype=torch.bfloat16)
    print("GEMM ALL", "OK" if allok else "FAIL")
    # quick timing
    M, N, K = 25600, 4096, 1024
    a = torch.randn(M, K, device=dev).bfloat16(); b = torch.randn(N, K, device=dev).bfloat16()
    for tn in (128, 256):
        c = ops.gemm_tn(a, b, tile_n=tn); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): ops.gemm_tn(a, b, out=c, tile_n=tn)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print("GEMM %dx%dx%d tile %d: %.3f ms  %.1f TFLOP/s" % (M, N, K, tn, ms, 2.0 * M * N * K / ms / 1e9))
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    c2 = a @ b.t(); torch.cuda.synchronize(); e0.record()
    for _ in range(10): c2 = a @ b.t()
    e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1) / 10
    print("cuBLAS bf16 same shape: %.3f ms %.1f TFLOP/s" % (ms, 2.0 * M * N * K / ms / 1e9))
