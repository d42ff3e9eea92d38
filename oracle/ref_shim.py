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

_HERE = os.path.dirname(os.path.abspath(__file__))
# the mounted reference in the build container; on the GPU box the byte-for-byte staged copy (oracle/build_ref.py)
_CANDIDATES = ["/root/reference/timit", os.path.join(_HERE, "_ref", "timit")]
REF_ROOT = next((c for c in _CANDIDATES if os.path.isfile(os.path.join(c, "models", "model_ctc.py"))), _CANDIDATES[0])


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "models", "model_ctc.py"))


def is_staged_copy():
    return REF_ROOT == _CANDIDATES[1]


def _install_editdistance_stub():
    if "editdistance" in sys.modules:
        return
    mod = types.ModuleType("editdistance")

    def _eval(a, b):
        # Levenshtein distance, row by row over the shorter sequence; the in-row dependency d[j] = min(c[j], d[j-1] + 1) is
        # a running minimum of c[j] - j (numpy), so a 600 x 60 problem costs 60 vector steps instead of 36 000 Python ones
        import numpy as np
        a, b = list(a), list(b)
        if len(a) < len(b):
            a, b = b, a
        if not b:
            return len(a)
        av = np.asarray([hash(x) for x in a], dtype=np.int64)
        idx = np.arange(len(a) + 1, dtype=np.int64)
        prev = idx.copy()
        for i, y in enumerate(b, 1):
            c = np.empty_like(prev)
            c[0] = i
            np.minimum(prev[1:] + 1, prev[:-1] + (av != hash(y)), out=c[1:])
            prev = np.minimum.accumulate(c - idx) + idx
        return int(prev[-1])

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


def load_train_loop():
    """The reference's own training loop body, `run_epoch` of steps/train_ctc.py:24-69, imported unmodified. Its module-level
    imports pull in utils/data_loader.py, which needs `kaldiio` (only used to read Kaldi ark files): stubbed."""
    load()
    sys.modules.setdefault("kaldiio", types.ModuleType("kaldiio"))
    import importlib
    mod = importlib.import_module("steps.train_ctc")
    return mod.run_epoch


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
