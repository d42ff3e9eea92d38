"""ORACLE — test infrastructure only. Generates tests/golden/* by running the UNMODIFIED reference
(/root/reference/timit, via oracle/ref_shim.py) plus the torch CPU kernels it calls, in the build container.

    python -m oracle.make_golden

The fixtures are small (inputs + outputs, weights re-created from a seed and verified by checksum) and are
what the `-m gpu` parity tests compare the CUDA path with on the GPU box, where /root/reference does not exist.
Reference call sites exercised: CTC_Model(...).forward (models/model_ctc.py:142-185), nn.CTCLoss(reduction='sum')
and loss/batch_size (steps/train_ctc.py:144,47-48), torch.max + compute_wer (train_ctc.py:51-52),
GreedyDecoder.decode / BeamDecoder.decode (utils/ctcDecoder.py:162-166,181-192; utils/BeamSearch.py:73-153).
"""
import json
import os
import sys

import numpy as np
import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_shim  # noqa: E402
from oracle.model_ref import synthetic_batch  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

MODEL_CASES = {
    # name: dict(T, N, F, H, L, C, bn, cnn, S, seed)
    "rnn_bn": dict(T=10, N=3, F=40, H=128, L=2, C=12, bn=True, cnn=False, S=4, seed=3),
    "rnn_nobn": dict(T=9, N=2, F=40, H=128, L=1, C=9, bn=False, cnn=False, S=3, seed=5),
    "cnn_rnn": dict(T=16, N=2, F=40, H=128, L=1, C=12, bn=True, cnn=True, S=3, seed=7),
    # LayerCNN's optional MaxPool2d (model_ctc.py:53-54), here over time after block 0. (The other entries of train_ctc.py:21's
    # supported_activate, tanh / sigmoid, cannot be constructed in the reference: `activation_function(inplace=True)`,
    # model_ctc.py:50, is a TypeError for them — the drop-in raises the same error from the same expression.)
    "cnn_pool": dict(T=24, N=2, F=40, H=128, L=1, C=12, bn=True, cnn=True, S=3, seed=9, act="relu",
                     cnn_layers=[[[1, 8], [3, 3], [1, 2], [1, 1], [2, 1]], [[8, 8], [3, 3], [1, 1], [1, 1], None]]),
}
CNN_LAYERS = [[(1, 32), (3, 3), (1, 2), (1, 1), None], [(32, 32), (3, 3), (2, 2), (1, 1), None]]


ACTS = {"relu": nn.ReLU, "tanh": nn.Tanh, "sigmoid": nn.Sigmoid}


def model_args(cfg):
    rnn_param = {"rnn_input_size": cfg["F"], "rnn_hidden_size": cfg["H"], "rnn_layers": cfg["L"],
                 "rnn_type": nn.LSTM, "bidirectional": True, "batch_norm": cfg["bn"]}
    layers = CNN_LAYERS
    if cfg.get("cnn_layers"):
        layers = [[tuple(l[0]), tuple(l[1]), tuple(l[2]), tuple(l[3]), tuple(l[4]) if l[4] else None] for l in cfg["cnn_layers"]]
    cnn_param = {"batch_norm": cfg["bn"], "activate_function": ACTS[cfg.get("act", "relu")], "layer": layers} if cfg["cnn"] else None
    return dict(add_cnn=cfg["cnn"], cnn_param=cnn_param, rnn_param=rnn_param, num_class=cfg["C"], drop_out=0.0)


def sample(t, limit=256):
    flat = t.detach().reshape(-1)
    step = max(1, flat.numel() // limit)
    return flat[::step][:limit].numpy().astype(np.float32), step


def gen_ctc():
    g = torch.Generator().manual_seed(11)
    T, N, C, S = 40, 6, 12, 8
    lp = torch.log_softmax(torch.randn(T, N, C, generator=g) * 2.0, -1)
    tl = torch.tensor([8, 5, 0, 6, 3, 8])
    tg = torch.zeros(N, S, dtype=torch.long)
    for n in range(N):
        tg[n, :tl[n]] = torch.randint(1, C, (int(tl[n]),), generator=g)
    tg[1, 1] = tg[1, 0]
    tg[1, 2] = tg[1, 0]  # repeated labels need blanks in between
    il = torch.tensor([40, 33, 21, 40, 2, 15])  # utterance 4: 3 labels in 2 frames -> infeasible (+inf)
    lpr = lp.clone().requires_grad_(True)
    nll = torch.nn.functional.ctc_loss(lpr, tg, il, tl, blank=0, reduction="none")
    feasible = torch.isfinite(nll)
    nll[feasible].sum().backward()
    loss_sum_ref = nn.CTCLoss(reduction="sum")(lp, tg, il, tl)  # what train_ctc.py:47 computes (inf here)
    np.savez_compressed(os.path.join(OUT, "ctc_small.npz"), log_probs=lp.numpy(), targets=tg.numpy(),
                        input_lengths=il.numpy(), target_lengths=tl.numpy(), nll=nll.detach().numpy(),
                        grad_feasible=lpr.grad.numpy(), feasible=feasible.numpy(), loss_sum=float(loss_sum_ref))
    print("ctc_small: nll", nll.detach().numpy())


def gen_models(ref, only=None):
    for name, cfg in MODEL_CASES.items():
        if only and name not in only:
            continue
        torch.manual_seed(cfg["seed"])
        model = ref.CTC_Model(**model_args(cfg))
        checksum = {k: float(v.double().abs().sum()) for k, v in model.state_dict().items()}
        x, frac, targets, tl = synthetic_batch(cfg["T"], cfg["N"], cfg["F"], cfg["C"], cfg["S"], cfg["seed"])
        model.train()
        out = model(x)
        T_out = out.shape[0]
        il = (frac * T_out).long()
        loss = nn.CTCLoss(reduction="sum")(out, targets, il, tl) / cfg["N"]
        loss.backward()
        _, index = torch.max(out, dim=-1)
        errs, toks = model.compute_wer(index.transpose(0, 1).numpy(), il.numpy(), targets.numpy(), tl.numpy())
        grads = {}
        for k, p in model.named_parameters():
            vals, step = sample(p.grad)
            grads[k] = dict(norm=float(p.grad.norm()), step=int(step), vals=vals)
        buffers = {k: v.detach().clone().numpy() for k, v in model.named_buffers() if "running" in k}
        model.eval()
        with torch.no_grad():
            out_eval = model(x)
        int2char = {i: ("blank" if i == 0 else "u%d" % i) for i in range(cfg["C"])}
        greedy = ref.GreedyDecoder(int2char, space_idx=-1, blank_index=0).decode(out_eval, il.tolist())
        payload = dict(x=x.numpy(), frac=frac.numpy(), targets=targets.numpy(), target_lengths=tl.numpy(),
                       input_lengths=il.numpy(), out_train=out.detach().numpy(), out_eval=out_eval.numpy(),
                       loss=float(loss), wer_errs=int(errs), wer_toks=int(toks), argmax=index.numpy())
        for k, g in grads.items():
            payload["gradvals/" + k] = g["vals"]
        for k, b in buffers.items():
            payload["buffer/" + k] = b
        np.savez_compressed(os.path.join(OUT, "model_%s.npz" % name), **payload)
        meta = dict(cfg=cfg, checksum=checksum, greedy=greedy,
                    grad_norm={k: g["norm"] for k, g in grads.items()}, grad_step={k: g["step"] for k, g in grads.items()})
        with open(os.path.join(OUT, "model_%s.json" % name), "w") as fh:
            json.dump(meta, fh, indent=1)
        print("model_%s: loss %.6f wer (%d,%d) greedy[0]=%r" % (name, float(loss), errs, toks, greedy[0]))


def gen_beam(ref):
    cases = []
    # small vocabulary, several widths / LM weights
    units8 = ["blank", "UNK", "a", "b", "c", "d", "e", "f"]
    arpa8 = ref_shim.write_synthetic_arpa(os.path.join(OUT, "lm_c8.arpa"), units8[1:], seed=0)
    units62 = ["blank", "UNK"] + ["p%02d" % i for i in range(60)]
    arpa62 = ref_shim.write_synthetic_arpa(os.path.join(OUT, "lm_c62.arpa"), units62[1:], seed=1)
    arrays = {}
    for tag, units, arpa, T, N, widths, alphas, seed in (
            ("c8", units8, arpa8, 30, 4, (3, 10), (0.01, 0.1), 21),
            ("c62", units62, arpa62, 50, 3, (20,), (0.1,), 22)):
        C = len(units)
        g = torch.Generator().manual_seed(seed)
        logits = 3.0 * torch.randn(T, N, C, generator=g)
        logits[:, :, 0] += 2.0
        lp = torch.log_softmax(logits, -1)
        lens = [T, T - 3, T // 2, T][:N]
        int2char = {i: u for i, u in enumerate(units)}
        probs = torch.exp(lp.transpose(0, 1)).contiguous()  # what BeamDecoder.decode hands to ctcBeamSearch.decode
        arrays["%s/log_probs" % tag] = lp.numpy()
        arrays["%s/probs" % tag] = probs.numpy()
        for w in widths:
            for a in alphas:
                dec = ref.BeamDecoder(int2char, beam_width=w, blank_index=0, space_idx=-1, lm_path=arpa, lm_alpha=a)
                strings = dec.decode(lp, lens)
                cases.append(dict(tag=tag, units=units, arpa=os.path.basename(arpa), T=T, N=N, lens=lens, beam_width=w,
                                  lm_alpha=a, strings=strings))
                print("beam %s w=%d a=%g -> %r" % (tag, w, a, strings[0][:40]))
    # an utterance that is blank everywhere: every frame is skipped, the empty prefix reaches the final LM step
    blank_err = None
    lp_blank = torch.log(torch.tensor([[[0.97, 0.01, 0.01, 0.01]]] * 6))  # [T=6, N=1, C=4]
    dec = ref.BeamDecoder({0: "blank", 1: "UNK", 2: "a", 3: "b"}, beam_width=3, blank_index=0, space_idx=-1,
                          lm_path=arpa8, lm_alpha=0.1)
    try:
        dec.decode(lp_blank, [6])
    except Exception as e:  # noqa: BLE001
        blank_err = type(e).__name__
    np.savez_compressed(os.path.join(OUT, "beam_small.npz"), **arrays)
    with open(os.path.join(OUT, "beam_small.json"), "w") as fh:
        json.dump(dict(cases=cases, all_blank_error=blank_err), fh, indent=1)
    print("all-blank utterance raises:", blank_err)


PACKED_CASE = dict(T=24, N=5, F=40, H=128, L=3, C=12, S=4, seed=11)


def gen_packed():
    """Variable-length (packed) semantics of the 863 path: the UNMODIFIED my_863_corpus/steps/model.py
    CTC_RNN on a pack_padded_sequence input, warp-ctc style loss restated with torch (oracle/packed_ref.warp_ctc_loss)."""
    sys.path.insert(0, "/root/reference/my_863_corpus/steps")
    import model as m863   # noqa: E402  (the reference's module, imported in place)
    from oracle import packed_ref
    cfg = PACKED_CASE
    torch.manual_seed(cfg["seed"])
    net = m863.CTC_RNN(rnn_input_size=cfg["F"], rnn_hidden_size=cfg["H"], rnn_layers=cfg["L"], rnn_type=nn.LSTM,
                       bidirectional=True, batch_norm=True, num_class=cfg["C"], drop_out=0.0)
    checksum = {k: float(v.double().abs().sum()) for k, v in net.state_dict().items()}
    x, lens, targets, tsz = packed_ref.synthetic_packed_batch(cfg["T"], cfg["N"], cfg["F"], cfg["C"], cfg["S"], cfg["seed"])
    net.train()
    act = net(nn.utils.rnn.pack_padded_sequence(x, lens))
    loss = packed_ref.warp_ctc_loss(act, targets, lens, tsz)
    loss.backward()
    grads = {}
    for k, p_ in net.named_parameters():
        vals, step = sample(p_.grad)
        grads[k] = dict(norm=float(p_.grad.norm()), step=int(step), vals=vals)
    buffers = {k: v.detach().clone().numpy() for k, v in net.named_buffers() if "running" in k}
    net.eval()
    with torch.no_grad():
        logp = net(nn.utils.rnn.pack_padded_sequence(x, lens))
    payload = dict(x=x.numpy(), lengths=np.asarray(lens, dtype=np.int64), targets=targets.numpy(),
                   target_sizes=np.asarray(tsz, dtype=np.int64), act_train=act.detach().numpy(), logp_eval=logp.numpy(),
                   loss=float(loss))
    for k, g in grads.items():
        payload["gradvals/" + k] = g["vals"]
    for k, b in buffers.items():
        payload["buffer/" + k] = b
    np.savez_compressed(os.path.join(OUT, "packed_rnn.npz"), **payload)
    with open(os.path.join(OUT, "packed_rnn.json"), "w") as fh:
        json.dump(dict(cfg=cfg, checksum=checksum, grad_norm={k: g["norm"] for k, g in grads.items()},
                       grad_step={k: g["step"] for k, g in grads.items()}), fh, indent=1)
    print("packed_rnn: loss %.6f, lengths %s" % (float(loss), lens))


BATCH_CASES = [(0, 0, 1, 1), (2, 2, 1, 1), (3, 1, 3, 1), (1, 2, 2, 4), (0, 0, 0, 3), (5, 5, 4, 2)]


def gen_batch():
    """Host-side batch assembly of the UNMODIFIED reference (utils/tools.py make_context / skip_feat, the n_downsample padding of
    SpeechDataset.__getitem__, utils/data_loader.py create_input; kaldiio stubbed, it is only used to read ark files)."""
    import types
    sys.path.insert(0, "/root/reference/timit")
    sys.path.insert(0, "/root/reference/timit/utils")
    sys.modules.setdefault("kaldiio", types.ModuleType("kaldiio"))
    import tools         # noqa: E402
    import data_loader   # noqa: E402
    rng = np.random.RandomState(21)
    feats = [rng.randn(L, 8).astype(np.float32) for L in (13, 1, 9, 20, 16)]
    labs = [rng.randint(1, 9, size=k).astype(np.int64) for k in (4, 1, 3, 6, 2)]
    payload = {"n": np.asarray(len(feats))}
    for i, (f, l) in enumerate(zip(feats, labs)):
        payload["feat/%d" % i] = f
        payload["label/%d" % i] = l
    for ci, (left, right, skip, down) in enumerate(BATCH_CASES):
        items = []
        for f, lab in zip(feats, labs):
            ft = tools.skip_feat(tools.make_context(f, left, right), skip)
            if ft.shape[0] % down != 0:
                ft = np.vstack([ft, np.zeros((down - ft.shape[0] % down, ft.shape[1]))])
            items.append((torch.from_numpy(ft), torch.LongTensor(lab), "utt"))
        x, isz, tg, tsz, _ = data_loader.create_input(items)
        payload["case%d/x" % ci] = x.numpy()
        payload["case%d/input_sizes" % ci] = isz.numpy()
        payload["case%d/targets" % ci] = tg.numpy()
        payload["case%d/target_sizes" % ci] = tsz.numpy()
    payload["cases"] = np.asarray(BATCH_CASES, dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, "batch_small.npz"), **payload)
    print("batch_small: %d cases" % len(BATCH_CASES))


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(4)
    ref = ref_shim.load()
    gen_ctc()
    gen_models(ref)
    gen_beam(ref)
    gen_packed()
    gen_batch()


if __name__ == "__main__":
    main()
