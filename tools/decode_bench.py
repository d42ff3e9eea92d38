"""cfg5: greedy and LM beam decode throughput at the TIMIT shape (T=800, N=32, C=62), next to the UNMODIFIED reference's
decoders (timit/utils/ctcDecoder.py:152-192, BeamSearch.py:73-153) timed on a bounded sample of the same batch; beam strings are
checked against the full-size golden fixture (all 32 utterances).

usage: python tools/decode_bench.py [beam_width=100] [cpu_utts=1] [out.json]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from ctc_pytorch_b200 import ops
from ctc_pytorch_b200.decoder import BeamDecoder, GreedyDecoder

beam_width = int(sys.argv[1]) if len(sys.argv) > 1 else 100
cpu_utts = int(sys.argv[2]) if len(sys.argv) > 2 else 1
out_path = sys.argv[3] if len(sys.argv) > 3 else None
T, N, C = 800, 32, 62
dev = "cuda"
units = ["blank", "UNK"] + ["p%02d" % i for i in range(60)]
int2char = dict(enumerate(units))
arpa = os.path.join(ROOT, "tests", "golden", "lm_c62.arpa")

# the posteriors of the full-size golden fixture (tests/golden/beam_full.json: strings of the UNMODIFIED reference's search for all
# 32 utterances at width 100, lm_alpha 0.1), built from integer arithmetic so the float32 bits are the same on every machine
from ctc_pytorch_b200 import synth
fixture = json.load(open(os.path.join(ROOT, "tests", "golden", "beam_full.json")))
alpha = fixture["cfg"]["lm_alpha"]
probs_np = synth.exact_probs(N, T, C, fixture["cfg"]["seed"])
lens = fixture["lens"]
log_probs = torch.from_numpy(probs_np).log().transpose(0, 1).contiguous()   # [T, N, C] on the host, as test_ctc.py:85 hands it over
unskipped = int(sum(((1.0 - probs_np[n, :l, 0]) >= np.float32(0.1)).sum() for n, l in enumerate(lens)))

res = {"shape": {"T": T, "N": N, "C": C, "beam_width": beam_width, "lm_alpha": alpha, "frames": int(sum(lens)),
                 "unskipped_frames": unskipped}}


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return r, best


lp_dev = log_probs.to(dev)
greedy = GreedyDecoder(int2char, space_idx=-1, blank_index=0)
beam = BeamDecoder(int2char, beam_width=beam_width, blank_index=0, space_idx=-1, lm_path=arpa, lm_alpha=alpha)

# device-resident input (kernel + string assembly) and end-to-end from the host tensor (H2D inside)
gs, t_g = timed(lambda: greedy.decode(lp_dev, lens), 5)
_, t_g_e2e = timed(lambda: greedy.decode(log_probs, lens), 5)
bs, t_b = timed(lambda: beam.decode(lp_dev, lens), 3)
_, t_b_e2e = timed(lambda: beam.decode(log_probs, lens), 3)
# kernel-only time of the search (CUDA events)
tab = beam._lm_table
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
ops.beam_search(lp_dev, lens, tab, beam_width, alpha, 0, input_is_log=True)
e1.record()
torch.cuda.synchronize()
t_b_dev = e0.elapsed_time(e1) * 1e-3
cands = unskipped * beam_width * (C - 1)
res["gpu"] = {"greedy_utt_s": N / t_g, "greedy_e2e_utt_s": N / t_g_e2e, "beam_utt_s": N / t_b, "beam_e2e_utt_s": N / t_b_e2e,
              "beam_search_call_ms": t_b_dev * 1e3, "beam_candidates_per_s": cands / t_b_dev,
              "greedy_ms": t_g * 1e3, "beam_ms": t_b * 1e3}

# kernel-level parity on ALL utterances: the same float32 probabilities the golden run of the unmodified reference consumed
labels = ops.beam_search(torch.from_numpy(probs_np).to(dev), lens, tab, beam_width, alpha, 0, input_is_log=False)
kernel_strings = [" ".join(units[l] for l in seq) for seq in labels]
res["parity"] = {"beam_strings_identical_to_reference_all_%d" % N: (kernel_strings == fixture["strings"]) if beam_width == fixture["cfg"]["beam_width"] else None}

# CPU: the UNMODIFIED reference decoders (pure Python, one thread) on `cpu_utts` utterances, via oracle/ref_shim.py
from oracle import ref_shim
if cpu_utts > 0 and ref_shim.available():
    ref = ref_shim.load()
    rg = ref.GreedyDecoder(int2char, space_idx=-1, blank_index=0)
    t0 = time.perf_counter()
    want_g = rg.decode(log_probs, lens)
    t_cg = time.perf_counter() - t0
    rb = ref.BeamDecoder(int2char, beam_width=beam_width, blank_index=0, space_idx=-1, lm_path=arpa, lm_alpha=alpha)
    t0 = time.perf_counter()
    want_bs = rb._decoder.decode(torch.from_numpy(probs_np[:cpu_utts]), lens[:cpu_utts])
    t_cb = time.perf_counter() - t0
    res["cpu"] = {"kind": "reference", "cores": 1, "greedy_utt_s": N / t_cg, "beam_utt_s": cpu_utts / t_cb,
                  "sample": "greedy: all %d utterances; beam: first %d utterance(s) (%d frames); unmodified ctcDecoder.py / BeamSearch.py"
                            % (N, cpu_utts, sum(lens[:cpu_utts]))}
    res["parity"]["greedy_strings_identical"] = gs == want_g
    res["parity"]["beam_strings_identical_on_cpu_sample"] = kernel_strings[:cpu_utts] == want_bs
    res["speedup"] = {"greedy": res["gpu"]["greedy_utt_s"] / res["cpu"]["greedy_utt_s"],
                      "beam": res["gpu"]["beam_utt_s"] / res["cpu"]["beam_utt_s"]}
print(json.dumps(res))
if out_path:
    json.dump(res, open(out_path, "w"), indent=1)
assert all(v is not False for v in res["parity"].values()), res["parity"]
