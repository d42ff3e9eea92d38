"""TEST INFRASTRUCTURE ONLY — restatement of the reference's host-side batch assembly: make_context / skip_feat
(timit/utils/tools.py:66-86), the n_downsample zero rows of SpeechDataset.__getitem__ (timit/utils/data_loader.py:103-110)
and create_input (data_loader.py:119-140). Index arithmetic instead of repeated vstack/hstack, so it is an independent
formulation; pinned against the live functions in tests/test_oracle.py."""
import numpy as np


def spliced(feature, left, right):
    """tools.py:66-75 — block c of the output row t is frame clamp(t + c - left, 0, L-1)."""
    feature = np.asarray(feature)
    L = feature.shape[0]
    idx = np.clip(np.arange(L)[:, None] + (np.arange(left + right + 1)[None, :] - left), 0, L - 1)   # [L, blocks]
    return feature[idx].reshape(L, -1)


def skipped(feature, skip):
    """tools.py:77-86 — keep frames 0, skip, 2 skip, ..."""
    if skip in (0, 1):
        return feature
    return feature[::skip]


def utterance(feature, left, right, skip, n_downsample):
    """data_loader.py:105-109."""
    feat = skipped(spliced(feature, left, right), skip)
    L = feat.shape[0]
    if L % n_downsample != 0:
        feat = np.vstack([feat, np.zeros((n_downsample - L % n_downsample, feat.shape[1]))])
    return feat


def batch(features, labels, left=0, right=0, skip=1, n_downsample=1):
    """create_input on the processed utterances: (inputs f32 [N,Tmax,F'], input_sizes f32, targets int64 [N,Smax], target_sizes int64)."""
    feats = [utterance(f, left, right, skip, n_downsample) for f in features]
    T_max = max(f.shape[0] for f in feats)
    S_max = max(len(l) for l in labels)
    N = len(feats)
    x = np.zeros((N, T_max, feats[0].shape[1]), dtype=np.float32)
    tg = np.zeros((N, S_max), dtype=np.int64)
    isz = np.zeros(N, dtype=np.float32)
    tsz = np.zeros(N, dtype=np.int64)
    for n, (f, l) in enumerate(zip(feats, labels)):
        x[n, :f.shape[0]] = f
        tg[n, :len(l)] = np.asarray(l, dtype=np.int64)
        isz[n] = np.float32(f.shape[0] / T_max)
        tsz[n] = len(l)
    return x, isz, tg, tsz
