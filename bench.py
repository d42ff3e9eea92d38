#!/usr/bin/env python
"""Benchmark of the CTC acoustic hot path (BASELINE.json: utterances/sec at T=800, N=32, feat=40, C=62).

    python bench.py --gpus 1 --steps K --warmup W            # our arm (sm_100a kernels through the C ABI)
    python bench.py --impl reference --steps K --warmup W    # the reference's own CPU path on the host cores

One "step" = one pass of the training hot path of timit/steps/train_ctc.py:44-65 over one synthetic batch of the
named shape (cfg2: 4 x BiLSTM-512 + BatchNorm, T=800, N=32 per GPU): forward, CTC loss / batch, frame arg-max +
collapse, backward, (N>1: one all-reduce of the flat gradient bucket), Adam step.
  value  utterances/s with the batch already resident in HBM;
  e2e    the same step driven through the public classes with HOST buffers: pinned H2D of features / labels /
         lengths and D2H of the loss and the collapsed arg-max labels inside the timed region.
Weak scaling: every rank processes its own N=32 shard. Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.nn as nn  # noqa: E402

CFG = {
    "cfg1": dict(T=800, N=4, F=40, C=62, H=256, L=2, S=60),
    "cfg2": dict(T=800, N=32, F=40, C=62, H=512, L=4, S=60),
    "cfg4": dict(T=1200, N=64, F=40, C=48, H=640, L=5, S=100),
}
# algorithmic FLOPs of the recurrent product per launch: 2 dirs * T * 2*4H*H*N


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler(object):
    """nvidia-smi clock / throttle sampling during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def make_batch(cfg, seed):
    from oracle.model_ref import synthetic_batch  # input generator only (seeded synthetic fbank batch)
    return synthetic_batch(cfg["T"], cfg["N"], cfg["F"], cfg["C"], cfg["S"], seed)


def rnn_param(cfg):
    return {"rnn_input_size": cfg["F"], "rnn_hidden_size": cfg["H"], "rnn_layers": cfg["L"], "rnn_type": nn.LSTM,
            "bidirectional": True, "batch_norm": True}


# --------------------------------------------------------------------------------------------------- reference arm
def cpu_threads():
    """Threads for the CPU arm. Measured on the GPU box's host (128 logical CPUs, profiles/cpu_thread_scaling_r1.txt):
    the reference's nn.LSTM step peaks at 16 intra-op threads (3.3 utt/s) and gets slower beyond (1.6 at 32, 0.6 at
    64), so 16 is "all the threads it can use"."""
    return max(1, min(16, os.cpu_count() or 1))


def cpu_reference_step_rate(cfg, n_utts, steps, warmup, threads):
    """The reference's CPU implementation of the path (nn.LSTM / BatchNorm1d / Linear / LogSoftmax / nn.CTCLoss /
    arg-max + collapse, composed as timit/models/model_ctc.py and train_ctc.py:44-65 compose them), restated in
    oracle/model_ref.py because /root/reference does not travel to the GPU box. Bounded sample: n_utts utterances."""
    from oracle.model_ref import RefAcousticModel
    from oracle import decode_ref
    torch.set_num_threads(threads)
    sub = dict(cfg, N=n_utts)
    torch.manual_seed(0)
    model = RefAcousticModel(cfg["F"], cfg["H"], cfg["L"], cfg["C"], batch_norm=True)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=0.005)
    loss_fn = nn.CTCLoss(reduction="sum")
    x, frac, tg, tl = make_batch(sub, 1)
    model.train()

    def step():
        out = model(x)
        il = (frac * out.shape[0]).long()
        loss = loss_fn(out, tg, il, tl) / n_utts
        _ = loss.item()
        idx = out.detach().argmax(-1).t().numpy()
        for n in range(n_utts):
            decode_ref.collapse(idx[n, :int(il[n])])
        opt.zero_grad()
        loss.backward()
        opt.step()
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return n_utts * steps / dt, dt / steps


def run_reference(args, cfg, rank, world):
    if rank != 0:
        return
    cores = cpu_threads()
    n_utts = 2
    warm = min(args.warmup, 1)
    rate, sec = cpu_reference_step_rate(cfg, n_utts, max(1, args.steps), warm, cores)
    line = {
        "impl": "reference", "metric": "utterances/sec (training step)", "value": rate, "unit": "utt/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": warm, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s: T=%d feat=%d C=%d %dxBiLSTM-%d+BN, CTC, greedy, Adam" % (
            args.config, cfg["T"], cfg["F"], cfg["C"], cfg["L"], cfg["H"]), "per_step_sample_utts": n_utts},
        "cpu_baseline": {"value": rate, "unit": "utt/s", "cores": cores, "kind": "port",
                         "sample": "%d utterances of the %s shape per step (torch CPU kernels, %d threads); the reference "
                                   "tree is not on the GPU box so its composition is restated in oracle/model_ref.py" % (
                                       n_utts, args.config, cores)},
        "e2e": {"value": rate, "unit": "utt/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# -------------------------------------------------------------------------------------------------------- our arm
def run_ours(args, cfg, rank, world):
    from ctc_pytorch_b200 import _lib, ops
    from ctc_pytorch_b200.model import CTC_Model
    from ctc_pytorch_b200.loss import CTCLoss
    from ctc_pytorch_b200.dist import GradBucket
    import torch.distributed as dist

    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    L = _lib.lib()
    T, N, F, C, H, Lyr = cfg["T"], cfg["N"], cfg["F"], cfg["C"], cfg["H"], cfg["L"]

    torch.manual_seed(0)
    model = CTC_Model(rnn_param=rnn_param(cfg), num_class=C, drop_out=0.0).to(dev)
    # the reference's optimizer (train_ctc.py:145: Adam + L2 weight decay), torch's single-launch (fused) implementation
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=0.005, fused=True)
    loss_fn = CTCLoss(reduction="sum")
    bucket = GradBucket(model.parameters())
    model.train()

    # several distinct host batches (pinned); the device-resident loop cycles over copies in HBM
    n_batches = 4
    host = []
    for b in range(n_batches):
        x, frac, tg, tl = make_batch(cfg, 100 * rank + b)
        host.append(tuple(t.pin_memory() for t in (x, frac, tg, tl)))
    devb = [tuple(t.to(dev) for t in hb) for hb in host]
    h2d_bytes = sum(t.numel() * t.element_size() for t in host[0])
    # L2 flush buffer (larger than the 126 MB L2) written between timed steps of the device-resident loop
    flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def step_dev(b, fetch=False, comm=True):
        x, frac, tg, tl = b
        out = model(x)
        out_len, bsz, _ = out.size()
        il = (frac * out_len).long()
        loss = loss_fn(out, tg, il, tl) / bsz
        _, labels, lens = ops.greedy_decode(out, il, blank=0)
        bucket.attach()
        loss.backward()
        if comm:
            bucket.allreduce_mean()
        opt.step()
        if fetch:
            return loss.item(), labels.cpu(), lens.cpu()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for w in range(max(args.warmup, 3)):
        step_dev(devb[w % n_batches])
    barrier()

    # ---- timed region 1: inputs resident in HBM ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = L.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for k in range(args.steps):
        flush.zero_()  # L2 flush between steps (counted inside the timed region: ~0.03 ms)
        step_dev(devb[k % n_batches])
    e1.record()
    barrier()
    ms_dev = e0.elapsed_time(e1)
    launches = L.launches - launches0
    clocks = sampler.stop() if rank == 0 else None

    # ---- timed region 2: end to end from host buffers ----
    barrier()
    d2h_bytes = 4 + N * T * 4 + N * 4
    e0.record()
    for k in range(args.steps):
        hb = host[k % n_batches]
        b = tuple(t.to(dev, non_blocking=True) for t in hb)
        step_dev(b, fetch=True)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)

    t_dev = torch.tensor([ms_dev, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e = t_dev.tolist()

    if rank != 0:
        return
    # ---- per-kernel pass (rank 0, N=1 semantics): device time of every C-ABI call, CUDA events on the launch stream ----
    per_call = {}
    orig_call = L.call

    def timed_call(name, *a):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        r = orig_call(name, *a)
        e.record()
        per_call.setdefault(name, []).append((s, e))
        return r
    L.call = timed_call
    prof_steps = 3
    overlap_was = model.overlap_wgrad
    model.overlap_wgrad = False  # kernels timed one at a time on one stream (the timed steps above overlap the wgrad GEMMs)
    for k in range(prof_steps):
        flush.zero_()
        step_dev(devb[k % n_batches], comm=False)  # rank 0 only: no collective in this diagnostic pass
    torch.cuda.synchronize()
    L.call = orig_call
    model.overlap_wgrad = overlap_was
    kern_ms = {nm: sum(s.elapsed_time(e) for s, e in v) / prof_steps for nm, v in per_call.items()}
    kern_cnt = {nm: len(v) // prof_steps for nm, v in per_call.items()}
    step_ms_prof = sum(kern_ms.values())

    pk = peaks()
    # dominant kernels: the persistent recurrent kernels (tensor-core work, latency bound by the per-step hand-off)
    rec_flops_launch = 2.0 * T * 2 * 4 * H * H * N           # one layer, both directions
    fwd_ms = kern_ms.get("ctcb200_lstm_fwd", 0.0) / max(1, kern_cnt.get("ctcb200_lstm_fwd", 1))
    bwd_ms = kern_ms.get("ctcb200_lstm_bwd", 0.0) / max(1, kern_cnt.get("ctcb200_lstm_bwd", 1))
    dom_name, dom_ms = ("lstm_bwd_kernel", bwd_ms) if bwd_ms >= fwd_ms else ("lstm_fwd_kernel", fwd_ms)
    ach = rec_flops_launch / (dom_ms * 1e-3) / 1e12 if dom_ms > 0 else 0.0
    # What actually bounds the recurrent kernels is the SM-to-SM (DSMEM) fabric: per time step every CTA receives and sends
    # the whole operand image of its cluster (fwd: H*16*2 B of h_t; BPTT: the same of dG plus the fp16 gate partials).
    dsmem_out = H * 16 * 2 + (4 * 16 * 32 * 2 if dom_name == "lstm_bwd_kernel" else 0)
    sm_hz = 1965e6
    dsmem_bpc = 2.0 * dsmem_out / (dom_ms * 1e-3 / T * sm_hz) if dom_ms > 0 else 0.0
    # DRAM traffic per launch of the dominant kernel from the committed `ncu --set full` capture (profiles/prof_lstm_r1.txt:
    # dram__bytes_read.sum + dram__bytes_write.sum), known for the cfg2 shape only
    ncu_traffic = {("cfg2", "lstm_bwd_kernel"): 528.6e6 + 186.5e6, ("cfg2", "lstm_fwd_kernel"): 423.7e6 + 378.5e6}
    roofline = {"kernel": dom_name, "bound": "tensor", "achieved": ach, "peak": pk["tf_sust"], "unit": "TFLOP/s",
                "frac": ach / pk["tf_sust"], "traffic": ncu_traffic.get((args.config, dom_name)),
                "traffic_unit": "bytes per launch (ncu, profiles/prof_lstm_r1.txt)", "peak_source": pk["src"] + " (sustained bf16)",
                "avg_launch_ms": dom_ms, "us_per_timestep": dom_ms * 1e3 / T,
                "limiter": {"resource": "DSMEM fabric (cluster all-gather / reduce-scatter every time step)",
                            "bytes_per_timestep_per_sm_in_plus_out": 2 * dsmem_out, "achieved_B_per_clk_per_sm": dsmem_bpc,
                            "peak_B_per_clk_per_sm": 17.0, "frac": dsmem_bpc / 17.0,
                            "peak_source": "B300_MICROARCH.md: 17 B/clk bidirectional per SM (measured on sm_103a), SM clock 1965 MHz"},
                "note": "recurrence: T dependent steps per launch, each moving the cluster's operand image between all CTAs; "
                        "tensor-pipe fraction reported for the contract, the DSMEM fraction is the binding one (DESIGN.md 3.2)"}
    # secondary rooflines: all dense GEMM launches together, and the CTC alpha/beta sweep
    rows = T * N
    gemm_flops = 0.0
    for l in range(Lyr):
        I = F if l == 0 else 2 * H
        gemm_flops += 2.0 * rows * 8 * H * I * (3 if l > 0 else 2)   # Gx, dWih (+ dX for l>0)
        gemm_flops += 2.0 * rows * 8 * H * H                          # dWhh (both directions)
    gemm_flops += 3 * 2.0 * rows * 2 * H * C
    gemm_ms = kern_ms.get("ctcb200_gemm_tn_bf16", 0.0)
    ctc_bytes = 2.0 * T * N * C * 4
    ctc_ms = kern_ms.get("ctcb200_ctc_loss_fwd", 0.0) + kern_ms.get("ctcb200_ctc_loss_bwd", 0.0)
    extra = {
        "gemm_all": {"bound": "tensor", "achieved": gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms else None,
                     "peak": pk["tf_sust"], "unit": "TFLOP/s", "ms_per_step": gemm_ms},
        "ctc_alpha_beta": {"bound": "hbm", "achieved": ctc_bytes / (ctc_ms * 1e-3) / 1e9 if ctc_ms else None,
                           "peak": pk["hbm"], "unit": "GB/s", "ms_per_step": ctc_ms,
                           "note": "N=32 is latency-bound (T dependent steps per warp)"},
        "kernel_ms_per_step": {k: round(v, 4) for k, v in sorted(kern_ms.items(), key=lambda kv: -kv[1])},
        "kernel_share_of_step": {k: round(v / step_ms_prof, 4) for k, v in sorted(kern_ms.items(), key=lambda kv: -kv[1])},
    }
    for k in ("gemm_all", "ctc_alpha_beta"):
        if extra[k]["achieved"]:
            extra[k]["frac"] = extra[k]["achieved"] / extra[k]["peak"]

    # ---- CPU baseline on the host cores (bounded sample) ----
    cores = cpu_threads()
    cpu_rate, cpu_sec = cpu_reference_step_rate(cfg, 2, 2, 1, cores)

    total_utts = N * world * args.steps
    line = {
        "metric": "utterances/sec (training step)", "value": total_utts / (ms_dev * 1e-3), "unit": "utt/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_dev / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "%s: T=%d N=%d/GPU feat=%d C=%d %dxBiLSTM-%d+BN; fwd + CTC loss + arg-max/collapse + bwd"
                               "%s + Adam" % (args.config, T, N, F, C, Lyr, H, " + grad all-reduce" if world > 1 else ""),
                   "global_batch": N * world, "parallelism": "dp%d" % world,
                   "timing": "CUDA events, max over ranks, 192 MiB L2 flush write between timed steps",
                   "numerics": "bf16 tensor-core operands, fp32 accumulate/state/loss"},
        "e2e": {"value": total_utts / (ms_e2e * 1e-3), "unit": "utt/s", "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": d2h_bytes, "ms_per_step": ms_e2e / args.steps,
                "note": "pinned H2D of x/frac/targets/lengths + D2H of loss and collapsed arg-max labels each step"},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roofline,
        "rooflines_other": extra,
        "cpu_baseline": {"value": cpu_rate, "unit": "utt/s", "cores": cores, "kind": "port",
                         "sample": "2 utterances x 2 steps of the same shape, torch CPU kernels (%d threads); "
                                   "oracle/model_ref.py restates the reference's composition" % cores},
    }
    emit(line)


_REAL_STDOUT = None


def quiet_stdout():
    """Route fd 1 to stderr while the run is in progress (NCCL and friends print banners on stdout from C); the one
    JSON line goes to the real stdout through emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        os.write(1, data)
    else:
        os.write(_REAL_STDOUT, data)


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(CFG))
    args = ap.parse_args()
    cfg = CFG[args.config]
    wd = int(os.environ.get("BENCH_WATCHDOG", "0"))
    if wd > 0:  # debugging aid: dump every thread's stack if the run is still alive after `wd` seconds
        import faulthandler
        faulthandler.dump_traceback_later(wd, repeat=True, file=sys.stderr)
    from ctc_pytorch_b200.dist import init_from_env
    if args.impl == "reference":
        rank = int(os.environ.get("RANK", "0"))
        run_reference(args, cfg, rank, int(os.environ.get("WORLD_SIZE", "1")))
        return
    rank, world = init_from_env()
    run_ours(args, cfg, rank, world)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
