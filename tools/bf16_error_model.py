"""CPU experiment: how large is the gradient error that bf16 tensor-core OPERANDS alone imply for this model? The fp32 oracle
composition is re-run with every matmul operand (activations, weights, h_{t-1}, and in the backward pass dG) rounded to bf16 and
fp32 accumulation — the arithmetic contract of the CUDA path, no kernels involved — and compared with plain fp32.
Supports the tolerance stated in DESIGN.md §4 (parameter gradients 3e-2 relative L2, measured 0.4-1.2 % on the GPU)."""
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn

from oracle.model_ref import synthetic_batch


class _Q(torch.autograd.Function):
    """bf16 rounding of a matmul operand in the forward pass and of the incoming gradient in the backward pass."""

    @staticmethod
    def forward(ctx, x):
        return x.bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().float()


def q(x, on):
    return _Q.apply(x) if on else x


def lstm_dir(x, w_ih, w_hh, reverse, quant):
    T, N, _ = x.shape
    H = w_hh.shape[1]
    gx = q(x, quant) @ q(w_ih, quant).t()
    h = x.new_zeros(N, H)
    c = x.new_zeros(N, H)
    out = [None] * T
    for t in (range(T - 1, -1, -1) if reverse else range(T)):
        g = gx[t] + q(h, quant) @ q(w_hh, quant).t()
        i, f, gg, o = g.chunk(4, 1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * torch.tanh(c)
        out[t] = h
    return torch.stack(out, 0)


def forward(params, x, quant, L):
    h = x.transpose(0, 1)
    for l in range(L):
        if l > 0:
            T, N, C = h.shape
            flat = h.reshape(T * N, C)
            mean, var = flat.mean(0), flat.var(0, unbiased=False)
            h = ((flat - mean) * torch.rsqrt(var + 1e-5) * params["bn%d.w" % l] + params["bn%d.b" % l]).reshape(T, N, C)
        h = torch.cat([lstm_dir(h, params["l%d.wih" % l], params["l%d.whh" % l], False, quant),
                       lstm_dir(h, params["l%d.wih_r" % l], params["l%d.whh_r" % l], True, quant)], -1)
    T, N, C = h.shape
    flat = h.reshape(T * N, C)
    mean, var = flat.mean(0), flat.var(0, unbiased=False)
    flat = (flat - mean) * torch.rsqrt(var + 1e-5) * params["fc.bnw"] + params["fc.bnb"]
    logits = q(flat, quant) @ q(params["fc.w"], quant).t()
    return torch.log_softmax(logits, -1).reshape(T, N, -1)


def main():
    T, N, F, H, L, C, S = (int(v) for v in (sys.argv[1:8] if len(sys.argv) >= 8 else (200, 8, 40, 256, 3, 62, 20)))
    torch.manual_seed(0)
    torch.set_num_threads(8)
    k = 1.0 / H ** 0.5
    params = {}
    for l in range(L):
        I = F if l == 0 else 2 * H
        for nm, shp in (("wih", (4 * H, I)), ("whh", (4 * H, H)), ("wih_r", (4 * H, I)), ("whh_r", (4 * H, H))):
            params["l%d.%s" % (l, nm)] = (torch.rand(shp) * 2 - 1) * k
        if l > 0:
            params["bn%d.w" % l] = torch.ones(2 * H)
            params["bn%d.b" % l] = torch.zeros(2 * H)
    params["fc.bnw"], params["fc.bnb"] = torch.ones(2 * H), torch.zeros(2 * H)
    params["fc.w"] = (torch.rand(C, 2 * H) * 2 - 1) / (2 * H) ** 0.5
    x, frac, tg, tl = synthetic_batch(T, N, F, C, S, 1)
    il = (frac * T).long()
    res = {}
    for quant in (False, True):
        ps = {k_: v.clone().requires_grad_(True) for k_, v in params.items()}
        out = forward(ps, x, quant, L)
        loss = nn.CTCLoss(reduction="sum")(out, tg, il, tl) / N
        loss.backward()
        res[quant] = (float(loss), {k_: v.grad.clone() for k_, v in ps.items()})
    l0, g0 = res[False]
    l1, g1 = res[True]
    print("T=%d N=%d H=%d L=%d: loss fp32 %.6f, bf16-operand model %.6f (rel %.2e)" % (T, N, H, L, l0, l1, abs(l1 - l0) / abs(l0)))
    worst = 0.0
    for k_ in g0:
        r = float((g1[k_] - g0[k_]).norm() / g0[k_].norm())
        worst = max(worst, r)
        print("  %-10s grad rel L2 %.3e" % (k_, r))
    print("worst parameter-gradient deviation implied by bf16 operands alone: %.3e" % worst)


if __name__ == "__main__":
    main()
