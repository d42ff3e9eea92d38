"""Streamed input projection / streamed gate gradients A/B on one GPU: bench.py's training step (cfg2 / cfg3 / cfg4, bf16 / x3)
with whole GEMMs around the recurrent kernels (overlap_gx / overlap_dg off) and with the time-chunked launches under them. Same Job, same
batches, L2 flushed between steps; prints one JSON line per setting.
    python tools/gx_stream_ab.py [cfg2] [bf16] [steps]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    cfg = dict(bench.CFG[name])
    job = bench.Job(name, cfg, 0, 1, dev, prec, strong=False)
    flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device=dev)
    rows = []
    for label, on, chunks, dg in (("whole", False, 8, False), ("gx8", True, 8, False), ("gx8_dg4", True, 8, 4), ("gx8_dg8", True, 8, 8),
                                  ("gx8_dg16", True, 8, 16), ("gx16_dg16", True, 16, 16), ("whole_again", False, 8, False),
                                  ("gx8_dg8_again", True, 8, 8)):
        job.model.overlap_gx, job.model.gx_chunks = on, chunks
        job.model.overlap_dg, job.model.dg_chunks = bool(dg), (dg or 8)
        for w in range(3):
            job.step(job.devb[w % len(job.devb)])
        ms = bench.timed_loop(job, steps, flush, 1, e2e=False) / steps
        # forward only (inference): the projection is a larger share of the pass
        job.model.eval()
        with torch.no_grad():
            for w in range(2):
                job.model(job.devb[0][0])
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for k in range(steps):
                flush.zero_()
                job.model(job.devb[k % len(job.devb)][0])
            e1.record()
            torch.cuda.synchronize()
        job.model.train()
        row = {"config": name, "precision": prec, "setting": label, "gx_chunks": chunks if on else 1, "dg_chunks": dg or 1, "train_step_ms": round(ms, 4),
               "forward_only_ms": round(e0.elapsed_time(e1) / steps, 4)}
        rows.append(row)
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
