"""Seeded synthetic inputs of the shapes BASELINE.json names. Pure input generation — no model arithmetic —
shared by bench.py, the tests and the oracle's golden-vector scripts so that every side sees the same batch.

`synthetic_batch` reproduces what the reference's `create_input` hands the training loop (timit/utils/data_loader.py:119-140):
features [N, T, F] f32 zero-padded past each utterance's end, lengths as float32 fractions of T, int64 labels zero-padded 2-D.
`exact_probs` builds class posteriors from integer arithmetic only, so the float32 values are bit-identical on every machine
(no exp/log in the recipe): the decode-identity fixtures depend on that.
"""
import numpy as np
import torch

CNN_LAYERS = [[(1, 32), (3, 3), (1, 2), (1, 1), None], [(32, 32), (3, 3), (2, 2), (1, 1), None]]   # conf/ctc_config.yaml:32-38

CONFIGS = {
    # BASELINE.json configs[0..3]; S = longest label sequence of the synthetic batch
    "cfg1": dict(T=800, N=4, F=40, C=62, H=256, L=2, S=60, cnn=False),
    "cfg2": dict(T=800, N=32, F=40, C=62, H=512, L=4, S=60, cnn=False),
    "cfg3": dict(T=800, N=32, F=40, C=62, H=512, L=4, S=60, cnn=True),
    "cfg4": dict(T=1200, N=64, F=40, C=48, H=640, L=5, S=100, cnn=False),
}


def synthetic_batch(T, N, feat, num_class, max_target, seed):
    """x ~ N(0,1) zeroed past each utterance's end, lengths linspace(1.0 -> 0.6)*T, targets uniform in [1, C-1] with
    S_n ~ U{S/2..S}, zero padded. Returns (x [N,T,F] f32, frac [N] f32, targets [N,S] i64, target_lengths [N] i64)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(N, T, feat, generator=g)
    lens = torch.linspace(1.0, 0.6, N).mul(T).round().long().clamp(min=1)
    for n in range(N):
        x[n, lens[n]:] = 0.0
    frac = (lens.float() / T)
    tl = torch.randint(max(1, max_target // 2), max_target + 1, (N,), generator=g)
    targets = torch.zeros(N, max_target, dtype=torch.long)
    for n in range(N):
        targets[n, :tl[n]] = torch.randint(1, num_class, (int(tl[n]),), generator=g)
    return x, frac, targets, tl


def model_kwargs(cfg, rnn_type=None, batch_norm=True, drop_out=0.0):
    """Constructor arguments of CTC_Model (reference and drop-in alike) for a CONFIGS entry."""
    import torch.nn as nn
    rnn_in = cfg["F"]
    rnn_param = {"rnn_input_size": rnn_in, "rnn_hidden_size": cfg["H"], "rnn_layers": cfg["L"],
                 "rnn_type": rnn_type or nn.LSTM, "bidirectional": True, "batch_norm": batch_norm}
    cnn_param = {"batch_norm": batch_norm, "activate_function": nn.ReLU, "layer": CNN_LAYERS} if cfg.get("cnn") else None
    return dict(add_cnn=bool(cfg.get("cnn")), cnn_param=cnn_param, rnn_param=rnn_param, num_class=cfg["C"], drop_out=drop_out)


def exact_probs(N, T, C, seed, active=0.3, blank=0):
    """Class posteriors [N, T, C] float32 that look like a trained CTC model's (most frames blank-dominant, short runs of
    frames peaked on one label) built from integers only: weights w = 2^e * m, p = w / sum(w) in float64, rounded once to
    float32. RandomState's legacy stream, integer shifts, one IEEE division and one IEEE rounding: bit-reproducible."""
    rs = np.random.RandomState(seed)
    e = rs.randint(0, 14, size=(N, T, C)).astype(np.int64)
    m = rs.randint(64, 128, size=(N, T, C)).astype(np.int64)
    w = (np.int64(1) << e) * m
    for n in range(N):
        t = 0
        while t < T:
            if rs.randint(0, 1000) < int(active * 1000):
                run = int(rs.randint(2, 7))
                k = int(rs.randint(1, C))
                for tt in range(t, min(T, t + run)):
                    w[n, tt, k] = (np.int64(1) << int(rs.randint(19, 23))) * int(rs.randint(64, 128))
                    w[n, tt, blank] = (np.int64(1) << int(rs.randint(14, 21))) * int(rs.randint(64, 128))
                t += run
            else:
                run = int(rs.randint(1, 9))
                for tt in range(t, min(T, t + run)):
                    w[n, tt, blank] = (np.int64(1) << int(rs.randint(24, 28))) * int(rs.randint(64, 128))
                t += run
    p = w.astype(np.float64) / w.sum(-1, keepdims=True).astype(np.float64)
    return p.astype(np.float32)


def exact_logprobs(T, N, C, seed):
    """[T, N, C] float32 'log-probabilities' made of exactly representable values (-k/64), with deliberate exact ties:
    input for the arg-max / collapse identity tests (greedy decoding only orders values, it never exponentiates)."""
    rs = np.random.RandomState(seed)
    k = rs.randint(1, 1024, size=(T, N, C)).astype(np.float32)
    lp = -(k / np.float32(64.0))
    runs = rs.randint(0, C, size=(T // 3 + 1, N))
    for t in range(T):                       # runs of three frames share an arg-max: exercises repeat removal
        for n in range(N):
            lp[t, n, runs[t // 3, n]] = np.float32(-1.0 / 128.0) if rs.randint(0, 4) else lp[t, n, runs[t // 3, n]]
    tie = rs.randint(0, T, size=(64,))
    for i, t in enumerate(tie):              # exact two-way ties for the maximum: the first index must win
        n = i % N
        a, b = sorted(rs.choice(C, size=2, replace=False).tolist())
        lp[t, n, a] = lp[t, n, b] = np.float32(0.0)
    return lp
