"""ORACLE — test infrastructure only (never imported by the product path).

Imports the *unmodified* reference (Diamondfan/CTC_pytorch, mounted read-only at /root/reference) so that
the oracle restatements can be pinned against it and golden vectors can be generated from it. The single
missing import on the hot path is `editdistance` (timit/models/model_ctc.py:7); a Levenshtein stub with the
same `eval(a, b)` contract is registered in sys.modules before the import. /root/reference exists only in the
build container — callers must check `available()`; nothing that runs on the GPU box depends on this module.
"""
import os
import sys
import types

REF_ROOT = "/root/reference/timit"


def available():
    return os.path.isdir(REF_ROOT)


def _install_editdistance_stub():
    if "editdistance" in sys.modules:
        return
    mod = types.ModuleType("editdistance")

    def _eval(a, b):
        a, b = list(a), list(b)
        prev = list(range(len(b) + 1))
        for i in range(1, len(a) + 1):
            cur = [i] + [0] * len(b)
            for j in range(1, len(b) + 1):
                cur[j] = min(cur[j - 1] + 1, prev[j] + 1, prev[j - 1] + (0 if a[i - 1] == b[j - 1] else 1))
            prev = cur
        return prev[len(b)]

    mod.eval = _eval
    sys.modules["editdistance"] = mod


def load():
    """Returns a namespace with the reference's CTC_Model, GreedyDecoder, BeamDecoder, ctcBeamSearch, LanguageModel."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    _install_editdistance_stub()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import importlib
    ns = types.SimpleNamespace()
    model_ctc = importlib.import_module("models.model_ctc")
    dec = importlib.import_module("utils.ctcDecoder")
    bs = importlib.import_module("utils.BeamSearch")
    lm = importlib.import_module("utils.NgramLM")
    ns.CTC_Model = model_ctc.CTC_Model
    ns.GreedyDecoder = dec.GreedyDecoder
    ns.BeamDecoder = dec.BeamDecoder
    ns.Decoder = dec.Decoder
    ns.ctcBeamSearch = bs.ctcBeamSearch
    ns.LanguageModel = lm.LanguageModel
    return ns


def write_synthetic_arpa(path, units, seed=0, bigram_frac=0.3):
    """Tab-separated bigram ARPA over `units` (+ <s>, </s>, <unk>) as the reference's IRSTLM recipe emits
    (SURVEY.md §8d): every unit is a unigram, ~30 % of the pairs get an explicit bigram."""
    import random
    rng = random.Random(seed)
    words = ["<s>", "</s>", "<unk>"] + [u for u in units if u not in ("<s>", "</s>", "<unk>")]
    uni = []
    for w in words:
        uni.append((round(-rng.uniform(0.5, 3.0), 6), w, round(-rng.uniform(0.05, 1.0), 6)))
    bi = []
    for w1 in words:
        if w1 == "</s>":
            continue
        for w2 in words:
            if w2 == "<s>":
                continue
            if rng.random() < bigram_frac:
                bi.append((round(-rng.uniform(0.1, 3.5), 6), w1 + " " + w2))
    with open(path, "w") as fh:
        fh.write("\n\\data\\\nngram 1=%d\nngram 2=%d\n\n\\1-grams:\n" % (len(uni), len(bi)))
        for p, w, b in uni:
            fh.write("%.6f\t%s\t%.6f\n" % (p, w, b))
        fh.write("\n\\2-grams:\n")
        for p, k in bi:
            fh.write("%.6f\t%s\n" % (p, k))
        fh.write("\n\\end\\\n")
    return path
