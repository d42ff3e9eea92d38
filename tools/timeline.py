"""Kernel timeline of ONE bench step (what nsys would show; CUPTI through torch.profiler): every kernel / memcpy with its stream,
start and duration, plus the idle gaps of the critical stream. Used to see what the step's critical path is made of and what
the NCCL all-reduces overlap with.

  python tools/timeline.py [cfg2] [bf16|x3] [out.json]                      # one GPU
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/timeline.py cfg2 bf16 out.json
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
    out = sys.argv[3] if len(sys.argv) > 3 else "gpurun_out/timeline.json"
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    cfg = dict(bench.CFG[name])
    job = bench.Job(name, cfg, rank, world, dev, prec, strong=False)
    for w in range(4):
        job.step(job.devb[w % len(job.devb)])
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for w in range(2):
            job.step(job.devb[w % len(job.devb)])
        torch.cuda.synchronize()
    if rank != 0:
        return
    trace = out + ".chrome.json"
    prof.export_chrome_trace(trace)
    with open(trace) as f:
        tr = json.load(f)
    os.remove(trace)
    ev = []
    for e in tr.get("traceEvents", []):
        if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset"):
            a = e.get("args", {})
            if a.get("device", local) != local:
                continue
            ev.append({"name": e["name"][:70], "stream": a.get("stream"), "start_us": float(e["ts"]), "dur_us": float(e["dur"])})
    ev.sort(key=lambda r: r["start_us"])
    if not ev:
        print("no CUDA events captured")
        return
    # second step only: cut after the optimizer kernels (the last launches) of the first step
    idx = [i for i, r in enumerate(ev) if "multi_tensor_apply" in r["name"]]
    cut = idx[len(idx) // 2 - 1] + 1 if len(idx) >= 2 else 0
    step = ev[cut:]
    t0 = step[0]["start_us"]
    for r in step:
        r["start_us"] = round(r["start_us"] - t0, 1)
        r["dur_us"] = round(r["dur_us"], 1)
    span = max(r["start_us"] + r["dur_us"] for r in step)
    streams = {}
    for r in step:
        streams.setdefault(r["stream"], []).append(r)
    summary = {"config": name, "precision": prec, "world": world, "span_us": span, "streams": {}}
    for s, rows in streams.items():
        busy = sum(r["dur_us"] for r in rows)
        summary["streams"][str(s)] = {"kernels": len(rows), "busy_us": round(busy, 1)}
    # gaps on the busiest stream
    main_s = max(streams, key=lambda s: sum(r["dur_us"] for r in streams[s]))
    rows = streams[main_s]
    gaps = []
    for a, b in zip(rows, rows[1:]):
        g = b["start_us"] - (a["start_us"] + a["dur_us"])
        if g > 15:
            gaps.append({"after": a["name"][:40], "before": b["name"][:40], "at_us": a["start_us"] + a["dur_us"], "gap_us": round(g, 1)})
    summary["main_stream"] = str(main_s)
    summary["gaps_over_15us_on_main_stream"] = gaps
    summary["gap_total_us"] = round(sum(g["gap_us"] for g in gaps), 1)
    with open(out, "w") as f:
        json.dump({"summary": summary, "events": step}, f)
    print(json.dumps(summary)[:3000])
    for r in step:
        print("%9.1f %8.1f  s%-3s %s" % (r["start_us"], r["dur_us"], r["stream"], r["name"]))


if __name__ == "__main__":
    main()
