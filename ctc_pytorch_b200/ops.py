"""Thin tensor-in / tensor-out wrappers over the C ABI (no arithmetic happens here)."""
import torch

from . import _lib


@_lib.on_tensor_device
def gemm_tn(a, b, out=None, out_dtype=torch.float32, accumulate=False, a_koff=0, b_koff=0, k=None, tile_n=0,
            max_ctas=0):
    """C[M,N] (+)= A[M, a_koff:a_koff+K] @ B[N, b_koff:b_koff+K]^T on the tcgen05 tensor cores.

    a, b: bf16, 2-D, K contiguous (row pitch may exceed the logical width but must be a multiple of 8).
    """
    _lib.require_cuda(a, b)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16, "gemm_tn takes bf16 operands"
    assert a.dim() == 2 and b.dim() == 2 and a.stride(1) == 1 and b.stride(1) == 1
    M, N = a.shape[0], b.shape[0]
    if k is None:
        k = a.shape[1] - a_koff
    assert a_koff + k <= a.shape[1] and b_koff + k <= b.shape[1]
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype, device=a.device)
    assert out.shape[0] == M and out.shape[1] == N and out.stride(1) == 1
    assert out.dtype in (torch.float32, torch.bfloat16)
    _lib.lib().call("ctcb200_gemm_tn_bf16", _lib.ptr(a), a.stride(0), _lib.ptr(b), b.stride(0), _lib.ptr(out),
                    out.stride(0), M, N, k, a_koff, b_koff, 1 if out.dtype == torch.bfloat16 else 0,
                    1 if accumulate else 0, tile_n, max_ctas, _lib.stream())
    return out


def gemm_atb(a, b, out=None, accumulate=False, a_roff=0, b_roff=0, k=None, tile_n=0, max_ctas=0):
    """C[M,N] (+)= A[a_roff:a_roff+K, :]^T @ B[b_roff:b_roff+K, :] on the tcgen05 tensor cores (MN-major operands): the
    weight-gradient form, operands as the forward / BPTT kernels left them (rows = the T*N frames)."""
    _lib.require_cuda(a, b)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16, "gemm_atb takes bf16 operands"
    assert a.dim() == 2 and b.dim() == 2 and a.stride(1) == 1 and b.stride(1) == 1
    M, N = a.shape[1], b.shape[1]
    if k is None:
        k = a.shape[0] - a_roff
    assert a_roff + k <= a.shape[0] and b_roff + k <= b.shape[0]
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=a.device)
    assert out.shape[0] == M and out.shape[1] == N and out.stride(1) == 1 and out.dtype == torch.float32
    _lib.lib().call("ctcb200_gemm_atb_bf16", _lib.ptr(a), a.stride(0), _lib.ptr(b), b.stride(0), _lib.ptr(out), out.stride(0),
                    M, N, k, a_roff, b_roff, 1 if accumulate else 0, tile_n, max_ctas, _lib.stream())
    return out


@_lib.on_tensor_device
def argmax_nt(log_probs, want_max=False):
    """Frame arg-max of [T,N,C] log-probs -> int32 [N,T] (first index on ties), optionally the max values."""
    _lib.require_cuda(log_probs)
    lp = log_probs.detach().float().contiguous()
    T, N, C = lp.shape
    idx = torch.empty((N, T), dtype=torch.int32, device=lp.device)
    mx = torch.empty((N, T), dtype=torch.float32, device=lp.device) if want_max else None
    _lib.lib().call("ctcb200_argmax", _lib.ptr(lp), T, N, C, _lib.ptr(idx), _lib.ptr(mx), _lib.stream())
    return (idx, mx) if want_max else idx


@_lib.on_tensor_device
def greedy_decode(log_probs, lengths, blank=0):
    """Arg-max + CTC collapse. Returns (idx [N,T] int32, labels [N,T] int32, label_lengths [N] int32)."""
    _lib.require_cuda(log_probs)
    lp = log_probs.detach().float().contiguous()
    T, N, C = lp.shape
    if torch.is_tensor(lengths):
        lens = lengths.to(device=lp.device, dtype=torch.int64).contiguous()
    else:
        lens = torch.as_tensor(list(lengths), dtype=torch.int64, device=lp.device)
    idx = torch.empty((N, T), dtype=torch.int32, device=lp.device)
    labels = torch.zeros((N, T), dtype=torch.int32, device=lp.device)
    out_len = torch.empty((N,), dtype=torch.int32, device=lp.device)
    _lib.lib().call("ctcb200_greedy_decode", _lib.ptr(lp), _lib.ptr(lens), T, N, C, int(blank), _lib.ptr(idx),
                    _lib.ptr(labels), _lib.ptr(out_len), _lib.stream())
    return idx, labels, out_len


@_lib.on_tensor_device
def edit_distance(hyp, hyp_len, ref, ref_len):
    """Levenshtein distance per row between int32 hypotheses hyp [N, *] (lengths hyp_len) and int64 references ref [N, *]
    (lengths ref_len), all on the device; returns int32 [N]. Feeds on greedy_decode's (labels, label_lengths)."""
    _lib.require_cuda(hyp, ref)
    dev = hyp.device
    hyp = hyp.to(torch.int32).contiguous()
    hyp_len = hyp_len.to(device=dev, dtype=torch.int32).contiguous()
    ref = ref.to(device=dev, dtype=torch.int64)
    if ref.dim() == 1:
        ref = ref.view(1, -1)
    ref = ref.contiguous()
    ref_len = ref_len.to(device=dev, dtype=torch.int64).contiguous()
    N = hyp.shape[0]
    if ref.shape[0] != N or hyp_len.numel() != N or ref_len.numel() != N:
        raise ValueError("edit_distance: batch sizes differ")
    dist = torch.empty(N, dtype=torch.int32, device=dev)
    _lib.lib().call("ctcb200_edit_distance", _lib.ptr(hyp), hyp.stride(0), _lib.ptr(hyp_len), _lib.ptr(ref), ref.stride(0),
                    _lib.ptr(ref_len), N, int(ref.shape[1]), _lib.ptr(dist), _lib.stream())
    return dist


_BEAM_ERRORS = {1: (IndexError, "tuple index out of range (the empty prefix reached the final LM step)"),
                2: (ValueError, "math domain error (log of a zero probability)"),
                3: (KeyError, "unit missing from the language model")}


@_lib.on_tensor_device
def beam_search(tensor, lengths, lm_table, beam_width, lm_alpha, blank=0, input_is_log=True):
    """CTC prefix beam search with a dense bigram table. `tensor` is [T,N,C] log-probs (input_is_log) or
    [N,T,C] float32 probabilities. Returns a list of label lists; raises the exception the reference's
    ctcBeamSearch.decode would raise for the same utterance (IndexError / ValueError / KeyError)."""
    _lib.require_cuda(tensor)
    L = _lib.lib()
    dev = tensor.device
    if input_is_log:
        lp = tensor.detach().float().contiguous()
        T, N, C = lp.shape
        probs = torch.empty((N, T, C), dtype=torch.float32, device=dev)
        L.call("ctcb200_exp_transpose", _lib.ptr(lp), _lib.ptr(probs), T, N, C, _lib.stream())
    else:
        probs = tensor.detach().float().contiguous()
        N, T, C = probs.shape
    if torch.is_tensor(lengths):
        lens = lengths.to(device=dev, dtype=torch.int64).contiguous()
    else:
        lens = torch.as_tensor(list(lengths), dtype=torch.int64, device=dev)
    lm_table = lm_table.to(device=dev, dtype=torch.float64).contiguous()
    assert lm_table.shape == (C + 1, C + 1)
    ws = torch.empty(L.dll.ctcb200_beam_workspace_bytes(T, N, C, int(beam_width)), dtype=torch.uint8, device=dev)
    out = torch.zeros((N, T), dtype=torch.int32, device=dev)
    out_len = torch.zeros((N,), dtype=torch.int32, device=dev)
    status = torch.zeros((N,), dtype=torch.int32, device=dev)
    L.call("ctcb200_beam_search", _lib.ptr(probs), _lib.ptr(lens), _lib.ptr(lm_table), float(lm_alpha), T, N, C,
           int(beam_width), int(blank), _lib.ptr(ws), _lib.ptr(out), _lib.ptr(out_len), _lib.ptr(status), _lib.stream())
    st = status.cpu().tolist()
    for n, code in enumerate(st):  # the reference decodes utterances in order and stops at the first failure
        if code:
            exc, msg = _BEAM_ERRORS[code]
            raise exc(msg)
    out = out.cpu().numpy()
    out_len = out_len.cpu().numpy()
    return [out[n, :out_len[n]].tolist() for n in range(N)]
