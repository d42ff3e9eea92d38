"""Drop-in for timit/utils/NgramLM.py: ARPA bigram language model used by the beam decoder.

Same class name, constructor and `get_uni_prob` / `get_bi_prob` / `score_bg` behaviour on the *tab-separated* ARPA
the reference's IRSTLM recipe writes: log10 scores converted to natural log, `UNK` aliases `<unk>`, an empty unit
means sentence start (as w1) or sentence end (as w2), a missing bigram backs off to backoff(w1) + unigram(w2).
The loader itself (`parse_arpa`) is more tolerant than the reference's: space-separated entries and higher-order
sections are parsed too. `dense_table` flattens `get_bi_prob` over the model's units
into the [(C+1) x (C+1)] float64 table the beam-search kernel reads from HBM/L2.
"""
import math
import re

import numpy as np


_SECTION = re.compile(r"^\\(\d+)-grams:\s*$")


def parse_arpa(path):
    """{order: {"w1 ... wn": (log10 prob, log10 back-off)}} from an ARPA file.

    Accepts what the reference's loader accepts — the TAB-separated layout IRSTLM / KenLM write (`prob<TAB>w1 w2<TAB>backoff`,
    timit/utils/NgramLM.py:42-55) — and, beyond it, entries whose fields are separated by plain spaces
    (SRILM style), sections of any order (`\\3-grams:` is parsed into its own table instead of leaking into the bigram
    table as in NgramLM.py:39-55), the `\\data\\` / `\\end\\` markers, blank lines and Windows line ends. An entry of order n
    has n words; a trailing extra field is the back-off weight."""
    tables = {}
    order = 0
    with open(path, "r") as fh:
        for raw in fh:
            line = raw.rstrip("\r\n")
            m = _SECTION.match(line.strip())
            if m:
                order = int(m.group(1))
                tables.setdefault(order, {})
                continue
            stripped = line.strip()
            if not stripped or stripped.startswith("\\") or order == 0:
                if stripped == "\\end\\":
                    order = 0
                continue
            if "\t" in line:
                fields = [f for f in line.split("\t")]
                if len(fields) < 2:
                    continue
                words, rest = fields[1], fields[2:]
            else:
                toks = stripped.split()
                if len(toks) < 1 + order:
                    continue
                words, rest = " ".join(toks[1:1 + order]), toks[1 + order:]
                fields = toks
            try:
                logp = float(fields[0])
                backoff = float(rest[0]) if rest and rest[0].strip() else 0.0
            except ValueError:
                continue
            tables[order][words] = (logp, backoff)
    return tables


class LanguageModel(object):
    def __init__(self, arpa_file=None, n_gram=2, start="<s>", end="</s>", unk="<unk>"):
        self.n_gram = n_gram
        self.start = start
        self.end = end
        self.unk = unk
        self.scale = math.log(10)     # ARPA stores log10; the search adds natural logs (NgramLM.py:22)
        self.initngrams(arpa_file)

    def initngrams(self, fn):
        tables = parse_arpa(fn)
        ln10 = self.scale
        self.unigram = {w: [ln10 * p, ln10 * b] for w, (p, b) in tables.get(1, {}).items()}
        self.bigram = {w: [ln10 * p, ln10 * b] for w, (p, b) in tables.get(2, {}).items()}
        self.higher = {n: t for n, t in tables.items() if n > 2}   # kept for inspection; the decoder scores bigrams only
        # class 1 of the acoustic model is spelled `UNK` (data_loader.py:16) and scores as the LM's unknown word (NgramLM.py:58)
        self.unigram["UNK"] = self.unigram[self.unk]

    def get_uni_prob(self, wid):
        return self.unigram[wid][0]

    def get_bi_prob(self, w1, w2):
        """ln p(w2 | w1) with back-off; KeyError for a unit that is not a unigram, as in the reference."""
        if w1 == "":
            w1 = self.start
        if w2 == "":
            w2 = self.end
        key = w1 + " " + w2
        if key not in self.bigram:
            return self.unigram[w1][1] + self.unigram[w2][0]
        return self.bigram[key][0]

    def score_bg(self, sentence):
        val = 0.0
        words = sentence.strip().split()
        val += self.get_bi_prob(self.start, words[0])
        for i in range(len(words) - 1):
            val += self.get_bi_prob(words[i], words[i + 1])
        val += self.get_bi_prob(words[-1], self.end)
        return val

    def dense_table(self, classes):
        """float64 [(C+1), (C+1)]: entry [i, j] = get_bi_prob(unit_i, unit_j); index C stands for the empty
        unit (sentence start as a row, sentence end as a column). NaN marks pairs the reference would raise
        KeyError on; the decoder reports those instead of scoring them."""
        C = len(classes)
        tab = np.full((C + 1, C + 1), np.nan, dtype=np.float64)
        for i in range(C + 1):
            w1 = classes[i] if i < C else ""
            for j in range(C + 1):
                w2 = classes[j] if j < C else ""
                try:
                    tab[i, j] = self.get_bi_prob(w1, w2)
                except KeyError:
                    pass
        return tab
