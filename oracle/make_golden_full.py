"""ORACLE — test infrastructure only. Full-shape golden vectors from the UNMODIFIED reference (/root/reference/timit via
oracle/ref_shim.py), at the shapes BASELINE.json names. Slow (minutes of CPU), so separate from make_golden.py:

    python -m oracle.make_golden_full model cfg2|cfg3|cfg4      # one training step of the reference's CTC_Model + nn.CTCLoss
    python -m oracle.make_golden_full model64 cfg2|cfg3|cfg4    # the same step with the reference model in float64
    python -m oracle.make_golden_full beam <part> <nparts>      # reference ctcBeamSearch, beam 100 + bigram LM, T=800, N=32
    python -m oracle.make_golden_full beam_merge <nparts>
    python -m oracle.make_golden_full beam_edges                # width 200 and the 0.9 / 0.1 threshold rows
    python -m oracle.make_golden_full greedy                    # reference GreedyDecoder at T=800, N=32

Inputs and weights are NOT stored: they are re-created on the GPU box from the seed (ctc_pytorch_b200/synth.py; weights by
torch's default initialisation, verified by checksum). Stored: loss, per-utterance nll, 256 sampled gradient entries and the
norm of every parameter gradient, BatchNorm running statistics, frame arg-max rows, a few log-prob rows, decode strings.
Reference call sites: models/model_ctc.py:142-185 (forward), steps/train_ctc.py:44-52,62-63 (loss, arg-max, backward),
utils/ctcDecoder.py:152-192, utils/BeamSearch.py:73-153.
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_shim  # noqa: E402
from ctc_pytorch_b200 import synth  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
FULL_SEED = {"cfg1": 1, "cfg2": 1, "cfg3": 2, "cfg4": 3}
BEAM_FULL = dict(T=800, N=32, C=62, seed=5, beam_width=100, lm_alpha=0.1, arpa="lm_c62.arpa")
UNITS62 = ["blank", "UNK"] + ["p%02d" % i for i in range(60)]


def sample(t, limit=256):
    flat = t.detach().reshape(-1)
    step = max(1, flat.numel() // limit)
    return flat[::step][:limit].numpy().astype(np.float32), step


def gen_model(name):
    ref = ref_shim.load()
    cfg = synth.CONFIGS[name]
    seed = FULL_SEED[name]
    torch.manual_seed(seed)
    model = ref.CTC_Model(**synth.model_kwargs(cfg))
    checksum = {k: float(v.double().abs().sum()) for k, v in model.state_dict().items()}
    x, frac, targets, tl = synth.synthetic_batch(cfg["T"], cfg["N"], cfg["F"], cfg["C"], cfg["S"], seed)
    model.train()
    t0 = time.time()
    out = model(x)
    T_out = out.shape[0]
    il = (frac * T_out).long()
    nll = nn.functional.ctc_loss(out, targets, il, tl, blank=0, reduction="none")
    loss = nn.CTCLoss(reduction="sum")(out, targets, il, tl) / cfg["N"]
    loss.backward()
    print("%s: reference step %.1f s, loss %.6f" % (name, time.time() - t0, float(loss)), flush=True)
    _, index = torch.max(out, dim=-1)
    errs, toks = model.compute_wer(index.transpose(0, 1).numpy(), il.numpy(), targets.numpy(), tl.numpy())
    payload = dict(input_lengths=il.numpy(), nll=nll.detach().numpy().astype(np.float64), loss=float(loss),
                   argmax=index.numpy().astype(np.uint8), wer_errs=int(errs), wer_toks=int(toks),
                   out_rows=np.arange(0, T_out, max(1, T_out // 16), dtype=np.int64))
    payload["out_sample"] = out.detach()[payload["out_rows"]].numpy()
    # margin between the best and second-best class of every frame: frames whose margin is below the stated log-prob tolerance
    # may legitimately flip their arg-max under bf16 operands
    top2 = torch.topk(out.detach(), 2, dim=-1).values
    payload["argmax_margin"] = (top2[..., 0] - top2[..., 1]).numpy().astype(np.float16)
    meta = dict(cfg=cfg, seed=seed, checksum=checksum, grad_norm={}, grad_step={})
    for k, p in model.named_parameters():
        vals, step = sample(p.grad)
        payload["gradvals/" + k] = vals
        meta["grad_norm"][k] = float(p.grad.norm())
        meta["grad_step"][k] = int(step)
    for k, b in model.named_buffers():
        if "running" in k:
            payload["buffer/" + k] = b.detach().clone().numpy()
    np.savez_compressed(os.path.join(OUT, "full_%s.npz" % name), **payload)
    with open(os.path.join(OUT, "full_%s.json" % name), "w") as fh:
        json.dump(meta, fh, indent=1)
    print("%s: wer (%d, %d); written" % (name, errs, toks))


def gen_model_f64(name):
    """The same step with the reference model switched to float64 (`model.double()`, nothing else changed): the yardstick for
    the float32 numbers above. At T=800 x 4 layers the reference's OWN float32 run deviates from this by ~1e-3 in the
    recurrent-weight gradients (stored here as `f32_vs_f64/<param>`), i.e. north_star's 1e-3 is the noise floor of the
    reference's arithmetic itself; the x3 mode of the CUDA path is asserted against these float64 values."""
    ref = ref_shim.load()
    cfg = synth.CONFIGS[name]
    seed = FULL_SEED[name]
    meta = json.load(open(os.path.join(OUT, "full_%s.json" % name)))
    g32 = np.load(os.path.join(OUT, "full_%s.npz" % name))
    torch.manual_seed(seed)
    model = ref.CTC_Model(**synth.model_kwargs(cfg))
    x, frac, targets, tl = synth.synthetic_batch(cfg["T"], cfg["N"], cfg["F"], cfg["C"], cfg["S"], seed)
    model = model.double()
    model.train()
    t0 = time.time()
    out = model(x.double())
    il = (frac * out.shape[0]).long()
    nll = nn.functional.ctc_loss(out, targets, il, tl, blank=0, reduction="none")
    loss = nn.CTCLoss(reduction="sum")(out, targets, il, tl) / cfg["N"]
    loss.backward()
    print("%s: float64 reference step %.1f s, loss %.9f" % (name, time.time() - t0, float(loss.detach())), flush=True)
    payload = dict(loss=float(loss.detach()), nll=nll.detach().numpy(), out_sample=out.detach()[torch.from_numpy(g32["out_rows"])].numpy())
    noise = {}
    for k, p in model.named_parameters():
        step = meta["grad_step"][k]
        v64 = p.grad.reshape(-1)[::step][:256]
        payload["gradvals/" + k] = v64.numpy()
        payload["gradnorm/" + k] = np.float64(p.grad.norm())
        v32 = torch.from_numpy(g32["gradvals/" + k]).double()
        noise[k] = float((v32 - v64).norm() / (v64.norm() + 1e-300))
        payload["f32_vs_f64/" + k] = np.float64(noise[k])
    np.savez_compressed(os.path.join(OUT, "full_%s_f64.npz" % name), **payload)
    worst = max((v, k) for k, v in noise.items() if not k.endswith("conv.bias"))
    print("%s: the reference's own float32 run vs float64: worst gradient rel-L2 %.2e (%s)" % (name, worst[0], worst[1]))


def _beam_decoder(ref, width, alpha, arpa):
    int2char = dict(enumerate(UNITS62))
    return ref.BeamDecoder(int2char, beam_width=width, blank_index=0, space_idx=-1,
                           lm_path=os.path.join(OUT, arpa), lm_alpha=alpha)


def beam_lengths(T, N):
    return [int(v) for v in torch.linspace(1.0, 0.6, N).mul(T).round().long().tolist()]


def gen_beam_part(part, nparts):
    ref = ref_shim.load()
    c = BEAM_FULL
    probs = synth.exact_probs(c["N"], c["T"], c["C"], c["seed"])
    lens = beam_lengths(c["T"], c["N"])
    dec = _beam_decoder(ref, c["beam_width"], c["lm_alpha"], c["arpa"])
    res = {}
    for n in range(part, c["N"], nparts):
        t0 = time.time()
        # kernel-level boundary of SURVEY.md §7: the float32 probabilities the reference's search consumes (BeamSearch.py:73)
        s = dec._decoder.decode(torch.from_numpy(probs[n:n + 1]), [lens[n]])
        res[n] = s[0]
        unskipped = int(((1 - probs[n, :lens[n], 0]) >= np.float32(0.1)).sum())
        print("beam utt %d: %.1f s, %d unskipped frames, %d labels" % (n, time.time() - t0, unskipped, len(s[0].split())), flush=True)
    with open(os.path.join(OUT, "_beam_full_part%d.json" % part), "w") as fh:
        json.dump(res, fh)


def gen_beam_merge(nparts):
    c = BEAM_FULL
    strings = [None] * c["N"]
    for part in range(nparts):
        path = os.path.join(OUT, "_beam_full_part%d.json" % part)
        for k, v in json.load(open(path)).items():
            strings[int(k)] = v
        os.remove(path)
    assert all(s is not None for s in strings)
    probs = synth.exact_probs(c["N"], c["T"], c["C"], c["seed"])
    with open(os.path.join(OUT, "beam_full.json"), "w") as fh:
        json.dump(dict(cfg=c, units=UNITS62, lens=beam_lengths(c["T"], c["N"]), strings=strings,
                       probs_checksum=float(probs.astype(np.float64).sum()),
                       probs_xor=int(np.bitwise_xor.reduce(probs.view(np.uint32).reshape(-1)))), fh, indent=1)
    print("beam_full.json written")


def f32(x):
    return np.float32(x)


def gen_beam_edges():
    """BeamDecoder's own default width (200) on two full-length utterances, and rows whose blank probability sits within one
    float32 ulp of the two thresholds of the search (BeamSearch.py:63 `mat[t-1, blank] < 0.9`, :93 `(1 - mat[t, blank]) < 0.1`)."""
    ref = ref_shim.load()
    c = BEAM_FULL
    out = {}
    probs = synth.exact_probs(2, c["T"], c["C"], 77)
    dec = _beam_decoder(ref, 200, 0.01, c["arpa"])
    t0 = time.time()
    out["width200"] = dict(seed=77, N=2, T=c["T"], lens=[c["T"], 640], lm_alpha=0.01, beam_width=200,
                           strings=dec._decoder.decode(torch.from_numpy(probs), [c["T"], 640]))
    print("width 200: %.1f s" % (time.time() - t0), flush=True)
    # threshold rows: blank probability patterns around 0.9 (float32) on frames that follow / carry a repeated label
    nine = f32(0.9)
    one_minus = f32(1.0) - f32(0.1)            # 1 - p == 0.1f exactly when p == 1 - 0.1f (float32 arithmetic)
    cands = [np.nextafter(nine, f32(0)), nine, np.nextafter(nine, f32(1)), np.nextafter(one_minus, f32(0)), one_minus,
             np.nextafter(one_minus, f32(1)), f32(0.8999), f32(0.9001)]
    C8 = 8
    units8 = ["blank", "UNK", "a", "b", "c", "d", "e", "f"]
    rs = np.random.RandomState(123)
    T, N = 24, len(cands)
    mat = np.zeros((N, T, C8), dtype=np.float32)
    for n, pb in enumerate(cands):
        for t in range(T):
            w = rs.randint(1, 64, size=C8).astype(np.float64)
            lab = 2 + (t // 4) % 3                 # label runs of 4 frames: repeats meet the `p_{t-1}(blank) < 0.9` branch
            w[lab] += 200.0
            rest = w[1:] / w[1:].sum()
            b = float(pb) if (t % 2 == 1) else float(rs.randint(1, 30)) / 64.0
            mat[n, t, 0] = np.float32(b)
            mat[n, t, 1:] = (rest * (1.0 - b)).astype(np.float32)
    int2char = dict(enumerate(units8))
    dec8 = ref.BeamDecoder(int2char, beam_width=6, blank_index=0, space_idx=-1, lm_path=os.path.join(OUT, "lm_c8.arpa"), lm_alpha=0.1)
    strings = dec8._decoder.decode(torch.from_numpy(mat), [T] * N)
    skipped = [[bool((f32(1) - mat[n, t, 0]) < 0.1) for t in range(T)] for n in range(N)]
    out["thresholds"] = dict(units=units8, T=T, N=N, beam_width=6, lm_alpha=0.1, blank_values=[float(v) for v in cands],
                             blank_bits=[int(np.float32(v).view(np.uint32)) for v in cands], strings=strings,
                             frames_skipped=[int(sum(r)) for r in skipped])
    np.savez_compressed(os.path.join(OUT, "beam_edges.npz"), thresholds=mat)
    with open(os.path.join(OUT, "beam_edges.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print("beam_edges: threshold strings", strings, "skipped frames", out["thresholds"]["frames_skipped"])


def gen_greedy():
    ref = ref_shim.load()
    T, N, C, seed = 800, 32, 62, 9
    lp = synth.exact_logprobs(T, N, C, seed)
    lens = beam_lengths(T, N)
    int2char = dict(enumerate(UNITS62))
    strings = ref.GreedyDecoder(int2char, space_idx=-1, blank_index=0).decode(torch.from_numpy(lp), lens)
    with open(os.path.join(OUT, "greedy_full.json"), "w") as fh:
        json.dump(dict(T=T, N=N, C=C, seed=seed, lens=lens, strings=strings,
                       lp_xor=int(np.bitwise_xor.reduce(lp.view(np.uint32).reshape(-1)))), fh, indent=1)
    print("greedy_full: %d strings, first %r" % (len(strings), strings[0][:40]))


def main():
    os.makedirs(OUT, exist_ok=True)
    what = sys.argv[1]
    torch.set_num_threads(int(os.environ.get("GOLDEN_THREADS", "4")))
    if what == "model":
        gen_model(sys.argv[2])
    elif what == "model64":
        gen_model_f64(sys.argv[2])
    elif what == "beam":
        gen_beam_part(int(sys.argv[2]), int(sys.argv[3]))
    elif what == "beam_merge":
        gen_beam_merge(int(sys.argv[2]))
    elif what == "beam_edges":
        gen_beam_edges()
    elif what == "greedy":
        gen_greedy()
    else:
        raise SystemExit("unknown target %r" % what)


if __name__ == "__main__":
    main()
