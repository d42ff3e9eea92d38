"""ORACLE — test infrastructure only (never imported by the product path).

Pure-Python restatement of the reference's decoding arithmetic, which (unlike the model and the loss)
is in-tree Python and is followed exactly, including the order of floating-point operations:

  * frame arg-max + CTC collapse            timit/models/model_ctc.py:187-202 (compute_wer),
                                            timit/utils/ctcDecoder.py:66-116,152-166 (GreedyDecoder)
  * Levenshtein distance, cer / wer         timit/utils/ctcDecoder.py:118-149
  * ARPA bigram LM with back-off            timit/utils/NgramLM.py:11-78
  * CTC prefix beam search with bigram LM   timit/utils/BeamSearch.py:9-153

Numeric conventions that matter for bit-identity: scores are Python floats (float64)
built from math.log of *float32* probabilities; LOG_ZERO = -99999999.0 is an absorbing sentinel in
log-add; the two thresholds ((1-p_blank) < 0.1 to skip a frame, p_{t-1}(blank) < 0.9 to pick the
blank-ending score for a repeated label) are evaluated in float32 (NumPy >= 2 weak-scalar rules);
beams are ranked by a stable descending sort, ties keeping first-insertion order; the final score
is the log score divided by the label count.

Pinned by tests/test_oracle.py against the reference's own classes imported from
/root/reference/timit (when present) and against golden vectors in tests/golden/.
"""
import math

import numpy as np

LOG_ZERO = -99999999.0


# ------------------------------------------------------------------------------------------------
# greedy path
# ------------------------------------------------------------------------------------------------
def collapse(frame_labels, blank=0):
    """Keep a frame's label iff it is not blank and differs from the previous frame's label."""
    out = []
    prev = None
    for j, v in enumerate(frame_labels):
        v = int(v)
        if v != blank and (j == 0 or v != prev):
            out.append(v)
        prev = v
    return out


def greedy_labels(log_probs, lengths, blank=0):
    """log_probs [T,N,C] array-like -> list of collapsed label lists (first index wins arg-max ties)."""
    lp = np.asarray(log_probs)
    idx = lp.argmax(axis=-1).T  # [N,T]
    return [collapse(idx[n, :int(lengths[n])], blank) for n in range(idx.shape[0])], idx


def greedy_strings(log_probs, lengths, int2char, blank=0):
    """What GreedyDecoder(space_idx=-1).decode returns: ' ' + char for every kept frame."""
    lp = np.asarray(log_probs)
    idx = lp.argmax(axis=-1).T
    blank_char = int2char[blank]
    res = []
    for n in range(idx.shape[0]):
        chars = [int2char[int(i)] for i in idx[n, :int(lengths[n])]]
        s = ""
        for i, ch in enumerate(chars):
            if ch == blank_char:
                continue
            if i != 0 and ch == chars[i - 1]:
                continue
            s += " " + ch
        res.append(s)
    return res


def levenshtein(a, b):
    """Edit distance between two sequences (unit costs)."""
    la, lb = len(a), len(b)
    if la == 0:
        return lb
    if lb == 0:
        return la
    row = list(range(lb + 1))
    for i in range(1, la + 1):
        diag = row[0]
        row[0] = i
        for j in range(1, lb + 1):
            keep = row[j]
            row[j] = min(row[j - 1] + 1, row[j] + 1, diag + (0 if a[i - 1] == b[j - 1] else 1))
            diag = keep
    return row[lb]


def cer(s1, s2):
    return levenshtein(s1, s2)


def wer(s1, s2):
    return levenshtein(s1.split(), s2.split())


def batch_errors(frame_idx, input_sizes, targets, target_sizes, blank=0):
    """(errors, tokens) as CTC_Model.compute_wer returns them."""
    errs = toks = 0
    for n in range(len(frame_idx)):
        ref = [int(v) for v in targets[n][:int(target_sizes[n])]]
        hyp = collapse(frame_idx[n][:int(input_sizes[n])], blank)
        errs += levenshtein(ref, hyp)
        toks += len(ref)
    return errs, toks


# ------------------------------------------------------------------------------------------------
# ARPA bigram language model
# ------------------------------------------------------------------------------------------------
class BigramLM(object):
    """Tab-separated ARPA file; log10 values converted to natural log by multiplying with ln(10)."""

    def __init__(self, arpa_path, start="<s>", end="</s>", unk="<unk>"):
        ln10 = math.log(10)
        self.start, self.end = start, end
        self.uni = {}
        self.bi = {}
        section = 0
        with open(arpa_path, "r") as fh:
            for raw in fh.readlines():
                line = raw.strip("\n")
                if line == "\\1-grams:":
                    section = 1
                    continue
                if line == "\\2-grams:":
                    section = 2
                    continue
                if section == 0:
                    continue
                parts = line.split("\t")
                table = self.uni if section == 1 else self.bi
                if len(parts) == 3:
                    table[parts[1]] = (ln10 * float(parts[0]), ln10 * float(parts[2]))
                elif len(parts) == 2:
                    table[parts[1]] = (ln10 * float(parts[0]), 0.0)
        self.uni["UNK"] = self.uni[unk]

    def bigram(self, w1, w2):
        if w1 == "":
            w1 = self.start
        if w2 == "":
            w2 = self.end
        hit = self.bi.get(w1 + " " + w2)
        if hit is not None:
            return hit[0]
        return self.uni[w1][1] + self.uni[w2][0]

    def table(self, classes):
        """Dense [(C+1) x (C+1)] float64 table: row = previous unit (last row = sentence start),
        column = next unit (last column = sentence end); NaN where the reference would raise KeyError."""
        C = len(classes)
        tab = np.full((C + 1, C + 1), np.nan)
        for i in range(C + 1):
            w1 = classes[i] if i < C else ""
            for j in range(C + 1):
                w2 = classes[j] if j < C else ""
                try:
                    tab[i, j] = self.bigram(w1, w2)
                except KeyError:
                    pass
        return tab


# ------------------------------------------------------------------------------------------------
# prefix beam search
# ------------------------------------------------------------------------------------------------
def log_add(x, y):
    if x <= LOG_ZERO:
        return y
    if y <= LOG_ZERO:
        return x
    if y - x > 0.0:
        x, y = y, x
    return x + math.log(1 + math.exp(y - x))


class _Hyp(object):
    __slots__ = ("total", "nonblank", "blank")

    def __init__(self):
        self.total = LOG_ZERO
        self.nonblank = LOG_ZERO
        self.blank = LOG_ZERO


def _ranked(hyps):
    """Prefixes by descending total score; Python's sort is stable so ties keep insertion order."""
    return [p for p, _ in sorted(hyps.items(), key=lambda kv: kv[1].total, reverse=True)]


def beam_search_one(probs, length, classes, beam_width, lm, lm_alpha, blank=0):
    """probs: float32 [T,C] probabilities of one utterance. Returns the best label tuple."""
    mat = np.asarray(probs, dtype=np.float32)
    C = mat.shape[1]
    one = np.float32(1.0)
    skip_thr = np.float32(0.1)
    rep_thr = np.float32(0.9)
    root = _Hyp()
    root.blank = 0.0
    root.total = 0.0
    last = {(): root}
    for t in range(int(length)):
        if (one - mat[t, blank]) < skip_thr:
            continue
        cur = {}
        prev_blank_lt = bool(mat[t - 1, blank] < rep_thr)
        log_blank = math.log(mat[t, blank])
        for y in _ranked(last)[:beam_width]:
            src = last[y]
            nb = LOG_ZERO
            if y:
                nb = src.nonblank + math.log(mat[t, y[-1]])
            bl = src.total + log_blank
            h = cur.get(y)
            if h is None:
                h = cur[y] = _Hyp()
            h.nonblank = log_add(h.nonblank, nb)
            h.blank = log_add(h.blank, bl)
            h.total = log_add(h.total, log_add(bl, nb))
            for k in range(C):
                if k == blank:
                    continue
                lm_term = 0.0
                if lm:
                    lm_term = lm.bigram(classes[y[-1]] if y else "", classes[k]) * lm_alpha
                base = src.blank if (y and y[-1] == k and prev_blank_lt) else src.total
                score = math.log(mat[t, k]) + lm_term + base
                ny = y + (k,)
                g = cur.get(ny)
                if g is None:
                    g = cur[ny] = _Hyp()
                g.nonblank = log_add(g.nonblank, score)
                g.total = log_add(g.total, score)
        last = cur
    final = {}
    for y in _ranked(last)[:beam_width]:
        eos = last[y].total + lm.bigram(classes[y[-1]], "") * lm_alpha  # IndexError for the empty prefix
        h = final.get(y)
        if h is None:
            h = final[y] = _Hyp()
        h.nonblank = log_add(h.nonblank, eos)
        h.total = log_add(h.total, eos)
    for y, h in final.items():
        n = len(y)
        h.total = h.total * (1.0 / (n if n else 1))
    return _ranked(final)[0]


def beam_search(probs, lengths, classes, beam_width, lm, lm_alpha, blank=0):
    """probs [N,T,C] float32 probabilities -> (list of label tuples, list of ' '-joined strings)."""
    labels, strings = [], []
    for n in range(len(probs)):
        best = beam_search_one(probs[n], lengths[n], classes, beam_width, lm, lm_alpha, blank)
        labels.append(tuple(int(v) for v in best))
        strings.append(" ".join(classes[l] for l in best))
    return labels, strings
