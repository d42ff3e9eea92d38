import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
print("cpu_count", os.cpu_count(), "torch threads", torch.get_num_threads(), flush=True)
import bench
cfg = bench.CFG["cfg2"]
for th in (8, 16, 32, 64):
    if th > (os.cpu_count() or 1): break
    t0 = time.time()
    rate, sec = bench.cpu_reference_step_rate(cfg, 2, 1, 1, th)
    print("threads %d: %.2f utt/s, %.2f s/step (wall %.1f)" % (th, rate, sec, time.time() - t0), flush=True)
