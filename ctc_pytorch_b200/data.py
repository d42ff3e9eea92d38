"""Device-side batch assembly: what SpeechDataset.__getitem__ + create_input do on the host with
NumPy (timit/utils/data_loader.py:99-140, timit/utils/tools.py:66-86) as one gather kernel over features that are already
on the GPU. Returns exactly create_input's first four outputs (inputs f32 [N, T_max, F'], input_sizes f32 fractions,
targets int64 [N, S_max], target_sizes int64), on the device."""
import torch

from . import _lib


def _lengths_after(L, skip, n_downsample):
    ls = L if skip <= 1 else (L + skip - 1) // skip
    if n_downsample > 1 and ls % n_downsample != 0:
        ls += n_downsample - ls % n_downsample
    return ls


def assemble_batch(features, labels, left_ctx=0, right_ctx=0, n_skip_frame=1, n_downsample=1, device="cuda"):
    """features: list of float tensors [L_n, F] (raw utterance features, host or device); labels: list of int sequences.
    One H2D copy of the concatenated features (if they are on the host), two kernels, no host-side NumPy work."""
    if not torch.cuda.is_available():
        raise RuntimeError("ctc_pytorch_b200.data.assemble_batch needs a CUDA device; there is no CPU path")
    N = len(features)
    if N == 0 or len(labels) != N:
        raise ValueError("assemble_batch: need the same, non-zero number of feature matrices and label sequences")
    F = int(features[0].shape[1])
    lens = [int(f.shape[0]) for f in features]
    if min(lens) <= 0:
        raise ValueError("assemble_batch: empty utterance")
    cat = torch.cat([torch.as_tensor(f, dtype=torch.float32) for f in features], 0).to(device, non_blocking=True).contiguous()
    offs = torch.zeros(N + 1, dtype=torch.int64)
    offs[1:] = torch.cumsum(torch.tensor(lens, dtype=torch.int64), 0)
    offs = offs.to(device)
    T_max = max(_lengths_after(L, n_skip_frame, n_downsample) for L in lens)
    Fo = F * (left_ctx + right_ctx + 1)
    x = torch.empty((N, T_max, Fo), dtype=torch.float32, device=device)
    input_sizes = torch.empty(N, dtype=torch.float32, device=device)
    L_ = _lib.lib()
    L_.call("ctcb200_assemble_features", _lib.ptr(cat), _lib.ptr(offs), N, F, int(left_ctx), int(right_ctx), int(n_skip_frame),
            int(n_downsample), T_max, _lib.ptr(x), _lib.ptr(input_sizes), _lib.stream())
    slens = [len(l) for l in labels]
    S_max = max(max(slens), 1)
    lab = torch.cat([torch.as_tensor(l, dtype=torch.int64).reshape(-1) for l in labels]).to(device)
    loffs = torch.zeros(N + 1, dtype=torch.int64)
    loffs[1:] = torch.cumsum(torch.tensor(slens, dtype=torch.int64), 0)
    loffs = loffs.to(device)
    targets = torch.empty((N, S_max), dtype=torch.int64, device=device)
    target_sizes = torch.empty(N, dtype=torch.int64, device=device)
    if lab.numel() == 0:
        lab = torch.zeros(1, dtype=torch.int64, device=device)
    L_.call("ctcb200_pad_labels", _lib.ptr(lab), _lib.ptr(loffs), N, S_max, _lib.ptr(targets), _lib.ptr(target_sizes), _lib.stream())
    return x, input_sizes, targets, target_sizes
