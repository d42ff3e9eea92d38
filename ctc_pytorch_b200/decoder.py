"""Drop-in for the reference's decoders, timit/utils/ctcDecoder.py (Decoder :9-149, GreedyDecoder
:152-166, BeamDecoder :168-192) with timit/utils/BeamSearch.py and timit/utils/NgramLM.py behind it.

Same constructors, `.decode(prob_tensor[T,N,C] log-probs, frame_seq_len) -> list[str]`, `.cer/.wer`,
`.num_word/.num_char` counters and string conventions (greedy strings start with a space, beam strings do
not). The arg-max / collapse and the prefix beam search run in libctcb200 kernels on the GPU; only the final
int -> string mapping and the edit-distance scoring of printed results stay on the host.
"""
import torch

from . import _lib
from . import ops
from .lm import LanguageModel


def collapse_frames(frame_labels, blank=0):
    """CTC collapse exactly as CTC_Model.compute_wer does it (timit/models/model_ctc.py:192-199): keep a
    frame's label iff it is not blank and differs from the previous frame's label."""
    out = []
    prev = None
    for j, v in enumerate(frame_labels):
        v = int(v)
        if v != blank and (j == 0 or v != prev):
            out.append(v)
        prev = v
    return out


def edit_distance(a, b):
    """Levenshtein distance (unit costs) between two sequences."""
    la, lb = len(a), len(b)
    if la == 0:
        return lb
    if lb == 0:
        return la
    row = list(range(lb + 1))
    for i in range(1, la + 1):
        diag, row[0] = row[0], i
        ai = a[i - 1]
        for j in range(1, lb + 1):
            keep = row[j]
            sub = diag if ai == b[j - 1] else diag + 1
            ins = row[j - 1] + 1
            dele = keep + 1
            row[j] = sub if sub < ins and sub < dele else (ins if ins < dele else dele)
            diag = keep
    return row[lb]


class Decoder(object):
    """Base class: turns network output into text so it can be scored against the labels."""

    def __init__(self, int2char, space_idx=1, blank_index=0):
        self.int_to_char = int2char
        self.space_idx = space_idx
        self.blank_index = blank_index
        self.num_word = 0
        self.num_char = 0

    def decode(self):
        raise NotImplementedError

    def phone_word_error(self, prob_tensor, frame_seq_len, targets, target_sizes):
        strings = self.decode(prob_tensor, frame_seq_len)
        targets = self._unflatten_targets(targets, target_sizes)
        target_strings = self._process_strings(self._convert_to_strings(targets))
        cer = 0
        wer = 0
        for x in range(len(target_strings)):
            cer += self.cer(strings[x], target_strings[x])
            wer += self.wer(strings[x], target_strings[x])
            self.num_word += len(target_strings[x].split())
            self.num_char += len(target_strings[x])
        return cer, wer

    def _unflatten_targets(self, targets, target_sizes):
        split_targets = []
        offset = 0
        for size in target_sizes:
            split_targets.append(targets[offset:offset + size])
            offset += size
        return split_targets

    def _process_strings(self, seqs, remove_rep=False):
        return [self._process_string(seq, remove_rep) for seq in seqs]

    def _process_string(self, seq, remove_rep=False):
        blank_char = self.int_to_char[self.blank_index]
        pieces = []
        for i, char in enumerate(seq):
            if char == blank_char:
                continue
            if remove_rep and i != 0 and char == seq[i - 1]:
                continue
            if self.space_idx == -1:
                pieces.append(" " + char)
            elif char == self.int_to_char[self.space_idx]:
                pieces.append(" ")
            else:
                pieces.append(char)
        return "".join(pieces)

    def _convert_to_strings(self, seq, sizes=None):
        strings = []
        for x in range(len(seq)):
            seq_len = sizes[x] if sizes is not None else len(seq[x])
            strings.append(self._convert_to_string(seq[x], seq_len))
        return strings

    def _convert_to_string(self, seq, sizes):
        result = [self.int_to_char[int(seq[i])] for i in range(sizes)]
        if self.space_idx == -1:
            return result
        return "".join(result)

    def wer(self, s1, s2):
        return edit_distance(s1.split(), s2.split())

    def cer(self, s1, s2):
        return edit_distance(s1, s2)

    _edit_distance = staticmethod(edit_distance)


def _to_device_logprobs(prob_tensor):
    """The reference hands the decoders a CPU tensor (test_ctc.py:85); accept that and any CUDA tensor.
    The decode itself always runs on the GPU."""
    if not torch.cuda.is_available():
        raise RuntimeError("ctc_pytorch_b200 decoders need a CUDA device; there is no CPU path")
    if not prob_tensor.is_cuda:
        prob_tensor = prob_tensor.cuda(non_blocking=True)
    return prob_tensor.detach().float().contiguous()


class GreedyDecoder(Decoder):
    """Per-frame arg-max, drop repeats, drop blanks (device kernels csrc/greedy.cu)."""

    def decode_labels(self, prob_tensor, frame_seq_len):
        """Integer form of decode(): list of collapsed label lists."""
        lp = _to_device_logprobs(prob_tensor)
        _, labels, lens = ops.greedy_decode(lp, frame_seq_len, blank=self.blank_index)
        labels = labels.cpu().numpy()
        lens = lens.cpu().numpy()
        return [labels[n, :lens[n]].tolist() for n in range(labels.shape[0])]

    def decode(self, prob_tensor, frame_seq_len):
        out = []
        for seq in self.decode_labels(prob_tensor, frame_seq_len):
            chars = [self.int_to_char[v] for v in seq]
            if self.space_idx == -1:
                out.append("".join(" " + c for c in chars))
            else:
                space = self.int_to_char[self.space_idx]
                out.append("".join(" " if c == space else c for c in chars))
        return out


class BeamDecoder(Decoder):
    """CTC prefix beam search with a bigram LM (device kernel csrc/beam.cu)."""

    def __init__(self, int2char, beam_width=200, blank_index=0, space_idx=-1, lm_path=None, lm_alpha=0.01):
        self.beam_width = beam_width
        super(BeamDecoder, self).__init__(int2char, space_idx=space_idx, blank_index=blank_index)
        self.lm_alpha = lm_alpha
        self.lm = LanguageModel(arpa_file=lm_path)
        self._lm_table = None  # uploaded lazily to the device the first decode runs on

    def _classes(self, C):
        return [self.int_to_char[i] for i in range(C)]

    def decode_labels(self, prob_tensor, frame_seq_len=None):
        lp = _to_device_logprobs(prob_tensor)
        T, N, C = lp.shape
        if frame_seq_len is None:
            frame_seq_len = [T] * N
        if self._lm_table is None or self._lm_table.device != lp.device or self._lm_table.shape[0] != C + 1:
            self._lm_table = torch.from_numpy(self.lm.dense_table(self._classes(C))).to(lp.device)
        return ops.beam_search(lp, frame_seq_len, self._lm_table, self.beam_width, self.lm_alpha, self.blank_index,
                               input_is_log=True)

    def decode(self, prob_tensor, frame_seq_len=None):
        return [" ".join(self.int_to_char[l] for l in seq) for seq in self.decode_labels(prob_tensor, frame_seq_len)]
