"""CTC alpha/beta + gradient kernels against the HBM roofline over the batch size: the TIMIT shape (T=800, C=62, S<=60) at
N = 32 ... 8192 utterances, CUDA-event time of the two C-ABI calls (ctcb200_ctc_loss_fwd / _bwd, no Python wrapper inside the
timed region), algorithmic bytes 2*T*N*C*4 (read the log-probs once, write the gradient once: SURVEY.md §8d) over that time.

usage: python tools/ctc_sweep_bench.py [out.json]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from ctc_pytorch_b200 import _lib

out_path = sys.argv[1] if len(sys.argv) > 1 else None
T, C, S = 800, 62, 60
dev = "cuda"
L = _lib.lib()
peak = 6582.5
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(pk):
    peak = json.load(open(pk))["hbm_gbs"]
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
rows = []
for N in (32, 128, 512, 2048, 4096, 8192):
    torch.manual_seed(N)
    lp = torch.randn(T, N, C, device=dev).log_softmax(-1)
    tl = torch.randint(S // 2, S + 1, (N,), device=dev)
    tg = torch.randint(1, C, (N, S), device=dev)
    il = torch.linspace(1.0, 0.6, N, device=dev).mul(T).round().long()
    ws = torch.empty(L.dll.ctcb200_ctc_workspace_floats(T, N, S), dtype=torch.float32, device=dev)
    nll = torch.empty(N, dtype=torch.float32, device=dev)
    grad = torch.empty_like(lp)
    st = _lib.stream()

    def fwd():
        L.call("ctcb200_ctc_loss_fwd", _lib.ptr(lp), _lib.ptr(tg), tg.stride(0), _lib.ptr(il), _lib.ptr(tl), T, N, C, S, 0,
               _lib.ptr(ws), _lib.ptr(nll), st)

    def bwd():
        L.call("ctcb200_ctc_loss_bwd", _lib.ptr(lp), _lib.ptr(tg), tg.stride(0), _lib.ptr(il), _lib.ptr(tl), T, N, C, S, 0,
               _lib.ptr(ws), _lib.ptr(nll), None, 1.0, _lib.ptr(grad), st)
    dll = L.dll
    default_thr = dll.ctcb200_ctc_set_fused_min_batch(-1)
    forms = {}
    for form, thr in (("latency", 1 << 30), ("throughput", 0)):
        dll.ctcb200_ctc_set_fused_min_batch(thr)
        for _ in range(3):
            fwd(); bwd()
        torch.cuda.synchronize()
        reps = 5
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        tf = tb = 0.0
        for _ in range(reps):
            flush.zero_()
            e[0].record(); fwd(); e[1].record(); bwd(); e[2].record()
            torch.cuda.synchronize()
            tf += e[0].elapsed_time(e[1]) / reps
            tb += e[1].elapsed_time(e[2]) / reps
        forms[form] = (tf, tb)
    dll.ctcb200_ctc_set_fused_min_batch(default_thr)
    used = "throughput" if N >= default_thr else "latency"
    tf, tb = forms[used]
    alg = 2.0 * T * N * C * 4
    valid = float(il.sum().item()) / (T * N)
    hrow = T * N * 4 * 32 * 4 * valid     # one history (KS=4 x 32 lanes, fp32) over the valid frames
    # latency form: alpha + beta histories written by the sweeps and read by the gradient kernel, log-probs read three times;
    # throughput form: the alpha history written once and read once, log-probs read twice
    traffic = {"latency": 4 * hrow + (3 * valid + 1) * alg / 2, "throughput": 2 * hrow + (2 * valid + 1) * alg / 2}
    row = {"N": N, "form": used, "fwd_ms": tf, "bwd_ms": tb, "algorithmic_MB": alg / 1e6,
           "achieved_GBs": alg / ((tf + tb) * 1e-3) / 1e9, "frac_of_hbm_peak": alg / ((tf + tb) * 1e-3) / 1e9 / peak,
           "valid_frame_fraction": valid, "expected_dram_traffic_MB": traffic[used] / 1e6,
           "expected_traffic_GBs": traffic[used] / ((tf + tb) * 1e-3) / 1e9, "utt_per_s": N / ((tf + tb) * 1e-3),
           "both_forms_ms": {k: {"fwd": v[0], "bwd": v[1], "total": v[0] + v[1]} for k, v in forms.items()}}
    rows.append(row)
    print(json.dumps(row))
    del lp, ws, grad
    torch.cuda.empty_cache()
res = {"shape": {"T": T, "C": C, "S_max": S}, "hbm_peak_GBs": peak, "rows": rows,
       "reading": "fwd = ctcb200_ctc_loss_fwd, bwd = ctcb200_ctc_loss_bwd. Latency form (small N): alpha and beta sweeps concurrently "
                  "in fwd (800 dependent steps per warp are the cost), parallel gradient kernel in bwd. Throughput form (N >= "
                  "ctcb200_ctc_set_fused_min_batch): alpha sweep in fwd, beta sweep fused with the gradient in bwd: one history "
                  "instead of two, log-probs read twice instead of three times. expected_dram_traffic counts what the form must "
                  "move through HBM (histories do not fit the 126 MB L2 at these N); see DESIGN.md 3.3"}
if out_path:
    json.dump(res, open(out_path, "w"), indent=1)
