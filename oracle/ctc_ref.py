"""ORACLE — test infrastructure only (never imported by the product path).

CPU restatement in numpy (float64 arithmetic on float32 inputs) of the CTC loss and gradient the
reference obtains from `nn.CTCLoss(reduction='sum')` (timit/steps/train_ctc.py:144,47,63).

The arithmetic itself lives in a third-party dependency that is not vendored under /root/reference:
PyTorch (requirements.txt pins nothing; README.md:1-2 says "pytorch1.2"; the torch installed here is
2.11.0). This file restates torch's published algorithm (aten/src/ATen/native/LossCTC.cpp:
ctc_loss_cpu_template / ctc_loss_backward_cpu_template — Graves et al. 2006, eqs. 6-8 and 16, in log
space) and is pinned by tests/test_oracle.py against torch's own CPU kernel and the committed golden
vectors generated through the reference's call site.

Conventions reproduced: blank index argument, 2-D zero-padded int64 targets, extended label sequence
l' of length 2S+1, +inf loss for infeasible alignments (zero_infinity=False), gradient w.r.t.
log-probs in torch's form exp(lp) - exp(logsum_{s: l'_s = c}(alpha_t(s)+beta_t(s)) + nll - lp),
zero rows for t >= input_length.
"""
import numpy as np

NEG_INF = -np.inf


def _logsumexp3(a, b, c):
    m = np.maximum(np.maximum(a, b), c)
    safe = np.where(np.isfinite(m), m, 0.0)
    with np.errstate(divide="ignore", invalid="ignore"):
        out = safe + np.log(np.exp(a - safe) + np.exp(b - safe) + np.exp(c - safe))
    return np.where(np.isfinite(m), out, NEG_INF)


def _extended(target, blank):
    S = len(target)
    ext = np.full(2 * S + 1, blank, dtype=np.int64)
    ext[1::2] = target
    skip = np.zeros(2 * S + 1, dtype=bool)  # skip[s]: transition s-2 -> s allowed
    if S > 1:
        skip[3::2] = ext[3::2] != ext[1:-2:2]
    return ext, skip


def ctc_alpha_beta(lp_n, target, T_n, blank=0):
    """One utterance. lp_n: [T, C] float array. Returns (nll, log_alpha [T_n, L], log_beta [T_n, L])."""
    ext, skip = _extended(np.asarray(target, dtype=np.int64), blank)
    L = len(ext)
    lp = np.asarray(lp_n, dtype=np.float64)
    la = np.full((T_n, L), NEG_INF)
    lb = np.full((T_n, L), NEG_INF)
    if T_n == 0:
        return (0.0 if L == 1 else np.inf), la, lb
    la[0, 0] = lp[0, blank]
    if L > 1:
        la[0, 1] = lp[0, ext[1]]
    for t in range(1, T_n):
        prev = la[t - 1]
        p1 = np.concatenate(([NEG_INF], prev))[:L]
        p2 = np.concatenate(([NEG_INF, NEG_INF], prev))[:L]
        p2 = np.where(skip, p2, NEG_INF)
        la[t] = _logsumexp3(prev, p1, p2) + lp[t, ext]
    tail = la[T_n - 1, L - 1]
    tail2 = la[T_n - 1, L - 2] if L > 1 else NEG_INF
    ll = _logsumexp3(np.array(tail), np.array(tail2), np.array(NEG_INF))
    nll = -float(ll)
    lb[T_n - 1, L - 1] = lp[T_n - 1, blank]
    if L > 1:
        lb[T_n - 1, L - 2] = lp[T_n - 1, ext[L - 2]]
    skip_out = np.concatenate((skip, [False, False]))[2:L + 2]  # transition s -> s+2 allowed
    for t in range(T_n - 2, -1, -1):
        nxt = lb[t + 1]
        n1 = np.concatenate((nxt, [NEG_INF]))[1:L + 1]
        n2 = np.concatenate((nxt, [NEG_INF, NEG_INF]))[2:L + 2]
        n2 = np.where(skip_out, n2, NEG_INF)
        lb[t] = _logsumexp3(nxt, n1, n2) + lp[t, ext]
    return nll, la, lb


def ctc_loss_and_grad(log_probs, targets, input_lengths, target_lengths, blank=0, grad_nll=None):
    """log_probs [T,N,C] float32; targets [N,S] int; lengths [N].

    Returns (nll [N] float64, grad [T,N,C] float64) with grad = d(sum_n grad_nll[n]*nll[n]) / d log_probs
    in torch's convention (see module docstring).
    """
    lp = np.asarray(log_probs)
    T, N, C = lp.shape
    nll = np.zeros(N)
    grad = np.zeros((T, N, C))
    for n in range(N):
        T_n = int(min(int(input_lengths[n]), T))
        S = int(target_lengths[n])
        tgt = np.asarray(targets[n][:S], dtype=np.int64)
        nll_n, la, lb = ctc_alpha_beta(lp[:, n, :], tgt, T_n, blank)
        nll[n] = nll_n
        g = 1.0 if grad_nll is None else float(grad_nll[n])
        ext, _ = _extended(tgt, blank)
        lpn = lp[:T_n, n, :].astype(np.float64)
        with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
            post = np.exp(la + lb + nll_n - lpn[:, ext])  # [T_n, L] state posteriors
            occ = np.zeros((T_n, C))
            for s, c in enumerate(ext):
                occ[:, c] += post[:, s]
            grad[:T_n, n, :] = (np.exp(lpn) - occ) * g
    return nll, grad
