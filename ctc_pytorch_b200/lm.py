"""Drop-in for timit/utils/NgramLM.py: ARPA bigram language model used by the beam decoder.

Same class name, constructor and `get_uni_prob` / `get_bi_prob` / `score_bg` behaviour: the file is the
*tab-separated* ARPA the reference's IRSTLM recipe writes, log10 scores are converted to natural log,
`UNK` aliases `<unk>`, an empty unit means sentence start (as w1) or sentence end (as w2), and a missing
bigram backs off to backoff(w1) + unigram(w2). `dense_table` flattens `get_bi_prob` over the model's units
into the [(C+1) x (C+1)] float64 table the beam-search kernel reads from HBM/L2.
"""
import math

import numpy as np


class LanguageModel(object):
    def __init__(self, arpa_file=None, n_gram=2, start="<s>", end="</s>", unk="<unk>"):
        self.n_gram = n_gram
        self.start = start
        self.end = end
        self.unk = unk
        self.scale = math.log(10)
        self.initngrams(arpa_file)

    def initngrams(self, fn):
        self.unigram = {}
        self.bigram = {}
        if self.n_gram == 3:
            self.trigrame = {}
        section = 0
        with open(fn, "r") as fh:
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
                fields = line.split("\t")
                table = self.unigram if section == 1 else self.bigram
                if len(fields) == 3:
                    table[fields[1]] = [self.scale * float(fields[0]), self.scale * float(fields[2])]
                elif len(fields) == 2:
                    table[fields[1]] = [self.scale * float(fields[0]), 0.0]
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
