"""TEST INFRASTRUCTURE ONLY — a torch-CPU emulation of the libctcb200 entry points the acoustic model drives
(ctc_pytorch_b200/model.py, cnn.py, packed.py), with the library's exact operand LAYOUTS (packed gate order, hi / lo bf16
parts, padded transposed operands, kernel-alignment of the packed path) and its arithmetic contract (bf16 operands, fp32
accumulation; three products in x3 mode).

Purpose: the host-side orchestration — which buffer goes where, in which layout and alignment, in every mode (bf16 / x3,
dropout, unidirectional, packed, frozen BatchNorm, CNN front with pooling) — can be checked against the oracle in the CPU
suite, without a GPU; the CUDA kernels themselves are checked by the `-m gpu` tests. Installed by `emulated()`, which patches
`_lib` so that "pointers" are the tensors themselves and every `call` lands here. Never imported by the product path.
"""
import contextlib

import torch
import torch.nn.functional as F


def _part(v, part):
    hi = v.to(torch.bfloat16)
    return (v - hi.float()).to(torch.bfloat16) if part else hi


def _orig_row(p, H):
    j, ul, q = p >> 7, (p & 127) >> 2, p & 3
    return q * H + j * 32 + ul


class Emu(object):
    def __init__(self):
        self.calls = []

    def call(self, name, *args):
        self.calls.append(name)
        fn = getattr(self, name[len("ctcb200_"):], None)
        if fn is None:
            raise NotImplementedError("emulator has no %s" % name)
        fn(*args)
        return 0

    # ---- layout kernels -------------------------------------------------------------------------------------------
    def pack_lstm_weights(self, wih_f, whh_f, wih_r, whh_r, wih_p, wihT_p, whh_p, whhT_p, H, I, Ipad, part, gates, stream):
        perm = torch.tensor([_orig_row(p, H) for p in range(4 * H)])

        def four(w):     # torch rows [gates*H, .] -> four gate slots, the unused ones zero
            w = w.detach()
            return torch.cat([w, torch.zeros((4 - gates) * H, w.shape[1])], 0) if gates < 4 else w
        for d, (wi, wh) in enumerate(((four(wih_f), four(whh_f)), (four(wih_r), four(whh_r)))):
            rows = slice(d * 4 * H, (d + 1) * 4 * H)
            wih_p[rows].zero_()
            wih_p[rows, :I] = _part(wi.detach()[perm], part)
            wihT_p[:, rows] = _part(wi.detach()[perm].t(), part)
            whh_p[rows] = _part(wh.detach()[perm], part)
            # whhT_p[(d, q, m), k] = W[q*H + k][m]
            whhT_p[rows] = _part(wh.detach().view(4, H, H).transpose(1, 2).reshape(4 * H, H), part)

    def cast_transpose(self, src, s_outer, s_inner, n_inner, scale, shift, dst, dst_pitch, dstT, dstT_pitch, n_pad, R, C, part,
                       stream):
        v = torch.as_strided(src.detach(), (R // n_inner, n_inner, C), (s_outer, s_inner, 1)).reshape(R, C).float()
        if scale is not None:
            v = v * scale + shift
        if dst is not None:
            dst[:, :C] = _part(v, part)
        if dstT is not None:
            cols = (torch.arange(R) // n_inner) * n_pad + torch.arange(R) % n_inner
            dstT[:, cols] = _part(v, part).t()

    def realign_rows(self, src, dst, lengths, T, N, W, split, direction, accumulate, stream):
        s = src.detach().reshape(T, N, W).clone()
        out = torch.zeros(T, N, W)
        for n in range(N):
            L = int(lengths[n])
            out[:L, n, :split] = s[:L, n, :split]
            if direction > 0:
                out[:L, n, split:] = s[T - L:, n, split:]
            else:
                out[T - L:, n, split:] = s[:L, n, split:]
        d = dst.view(T, N, W)
        if accumulate:
            d += out
        else:
            d.copy_(out)

    def dropout_apply(self, a, mask, inv_keep, n, stream):
        a.mul_(mask.to(a.dtype) * inv_keep)

    # ---- dense contraction ------------------------------------------------------------------------------------------
    def gemm_tn_bf16(self, A, lda, B, ldb, C, ldc, M, N, K, a_koff, b_koff, out_bf16, accumulate, tile_n, max_ctas, stream):
        assert A.dtype == torch.bfloat16 and B.dtype == torch.bfloat16
        assert A.stride(0) == lda and B.stride(0) == ldb and C.stride(0) == ldc
        prod = A[:M, a_koff:a_koff + K].float() @ B[:N, b_koff:b_koff + K].float().t()
        if accumulate:
            C[:M, :N] += prod.to(C.dtype)
        else:
            C[:M, :N] = prod.to(C.dtype)

    def gemm_atb_bf16(self, A, lda, B, ldb, C, ldc, M, N, K, a_roff, b_roff, accumulate, tile_n, max_ctas, stream):
        assert A.dtype == torch.bfloat16 and B.dtype == torch.bfloat16
        assert A.stride(0) == lda and B.stride(0) == ldb and C.stride(0) == ldc
        prod = A[a_roff:a_roff + K, :M].float().t() @ B[b_roff:b_roff + K, :N].float()
        if accumulate:
            C[:M, :N] += prod
        else:
            C[:M, :N] = prod

    # ---- recurrent kernels ------------------------------------------------------------------------------------------
    @staticmethod
    def _unpack_cols(H):
        """packed column (within one direction) -> (gate q, unit) as index tensors"""
        p = torch.arange(4 * H)
        return p & 3, (p >> 7) * 32 + ((p & 127) >> 2)

    def lstm_fwd(self, gx, whh, whh_lo, hout, c_save, gates, scratch, T, N, H, tile, cell, stream):
        x3 = whh_lo is not None
        q_of, u_of = self._unpack_cols(H)
        for d in range(2):
            W_hi = whh[d * 4 * H:(d + 1) * 4 * H].float()
            W_lo = whh_lo[d * 4 * H:(d + 1) * 4 * H].float() if x3 else None
            g = gx.view(T, N, 8 * H)[:, :, d * 4 * H:(d + 1) * 4 * H]
            h = torch.zeros(N, H)
            c = torch.zeros(N, H)
            for t in (range(T) if d == 0 else range(T - 1, -1, -1)):
                h_hi = h.to(torch.bfloat16).float()
                rec = h_hi @ W_hi.t()
                if x3:
                    h_lo = (h - h_hi).to(torch.bfloat16).float()
                    rec = rec + h_lo @ W_hi.t() + h_hi @ W_lo.t()
                pre = g[t] + rec                                  # packed columns
                gate = torch.empty(N, 4, H)
                gate[:, q_of, u_of] = pre
                if cell == 0:
                    i, f, gg, o = torch.sigmoid(gate[:, 0]), torch.sigmoid(gate[:, 1]), torch.tanh(gate[:, 2]), torch.sigmoid(gate[:, 3])
                    c = f * c + i * gg
                    h = o * torch.tanh(c)
                    saved = torch.stack([i, f, gg, o], -1)
                elif cell == 1:      # GRU: slots (r, z, n, -); the n gate takes r * (W_hn h) from the recurrent product alone
                    recg = torch.empty(N, 4, H)
                    recg[:, q_of, u_of] = rec
                    r_, z_ = torch.sigmoid(gate[:, 0]), torch.sigmoid(gate[:, 1])
                    hn = recg[:, 2]
                    nn_ = torch.tanh(gate[:, 2] - hn + r_ * hn)
                    h = (1 - z_) * nn_ + z_ * h
                    c = h
                    saved = torch.stack([r_, z_, nn_, hn], -1)
                else:                # vanilla RNN, slot 0
                    h = torch.relu(gate[:, 0]) if cell == 3 else torch.tanh(gate[:, 0])
                    c = h
                    saved = torch.stack([h, torch.zeros_like(h), torch.zeros_like(h), torch.zeros_like(h)], -1)
                rows = slice(t * N, (t + 1) * N)
                hout[rows, d * H:(d + 1) * H] = h
                if c_save is not None:
                    c_save[rows, d * H:(d + 1) * H] = c
                if gates is not None:
                    gates[rows, d * H:(d + 1) * H] = saved.to(gates.dtype)

    def lstm_bwd(self, dhout, whhT, whhT_lo, c_save, gates, dg, dg_lo, dg_rec, dg_rec_lo, scratch, T, N, H, tile, cell, bn_x,
                 bn_coef, res_counter, res_event, stream):
        x3 = whhT_lo is not None
        dh_all = dhout.detach().view(T, N, 2 * H).float()
        if bn_x is not None:
            coef = bn_coef.view(3, 2 * H)
            dh_all = coef[0] * dh_all + coef[1] * bn_x.view(T, N, 2 * H) + coef[2]
        q_of, u_of = self._unpack_cols(H)
        # packed dg column of (unit, gate): dir*4H + (unit>>5)*128 + (unit&31)*4 + gate
        for d in range(2):
            WT_hi = whhT[d * 4 * H:(d + 1) * 4 * H].float().view(4, H, H)      # [q, m, k]
            WT_lo = whhT_lo[d * 4 * H:(d + 1) * 4 * H].float().view(4, H, H) if x3 else None
            dc_carry = torch.zeros(N, H)
            dG_prev = None                                                      # [N, 4, H] of the previous BPTT step
            order = range(T - 1, -1, -1) if d == 0 else range(T)
            for t in order:
                rows = slice(t * N, (t + 1) * N)
                tprev = t - 1 if d == 0 else t + 1
                has_prev = 0 <= tprev < T
                dh = dh_all[t, :, d * H:(d + 1) * H].clone()
                if dG_prev is not None:
                    g_hi = dG_prev.to(torch.bfloat16).float()
                    rec = torch.einsum("qmk,nqk->nm", WT_hi, g_hi)
                    if x3:
                        g_lo = (dG_prev - g_hi).to(torch.bfloat16).float()
                        rec = rec + torch.einsum("qmk,nqk->nm", WT_hi, g_lo) + torch.einsum("qmk,nqk->nm", WT_lo, g_hi)
                    dh = dh + rec
                gt = gates[rows, d * H:(d + 1) * H].float()
                gi, gf, gg, go = gt[..., 0], gt[..., 1], gt[..., 2], gt[..., 3]
                c_t = c_save[rows, d * H:(d + 1) * H]
                c_p = c_save[tprev * N:(tprev + 1) * N, d * H:(d + 1) * H] if has_prev else torch.zeros(N, H)
                zero = torch.zeros_like(dh)
                if cell == 0:
                    tc = torch.tanh(c_t)
                    d_o = dh * tc * go * (1 - go)
                    dc = dc_carry + dh * go * (1 - tc * tc)
                    d_i = dc * gg * gi * (1 - gi)
                    d_f = dc * c_p * gf * (1 - gf)
                    d_g = dc * gi * (1 - gg * gg)
                    dc_carry = dc * gf
                    dG_prev = torch.stack([d_i, d_f, d_g, d_o], 1)              # [N, 4, H]: what W_hh^T multiplies next step
                    dG_in = dG_prev
                elif cell == 1:      # saved (r, z, n, hn), c_p = h_prev
                    dh = dh + dc_carry
                    dn_pre = dh * (1 - gf) * (1 - gg * gg)
                    d_r = dn_pre * go * gi * (1 - gi)
                    d_z = dh * (c_p - gg) * gf * (1 - gf)
                    dc_carry = dh * gf
                    dG_prev = torch.stack([d_r, d_z, dn_pre * gi, zero], 1)
                    dG_in = torch.stack([d_r, d_z, dn_pre, zero], 1)
                else:
                    dpre = dh * (gi > 0).float() if cell == 3 else dh * (1 - gi * gi)
                    dG_prev = torch.stack([dpre, zero, zero, zero], 1)
                    dG_in = dG_prev

                def store(dst, dst_lo, val):
                    packed = val[:, q_of, u_of]                                 # [N, 4H] packed columns of this direction
                    hi = packed.to(torch.bfloat16)
                    dst[rows, d * 4 * H:(d + 1) * 4 * H] = hi
                    if x3:
                        dst_lo[rows, d * 4 * H:(d + 1) * 4 * H] = (packed - hi.float()).to(torch.bfloat16)
                store(dg, dg_lo, dG_in)
                if cell == 1:
                    store(dg_rec, dg_rec_lo, dG_prev)

    # ---- BatchNorm / softmax --------------------------------------------------------------------------------------------
    def bn_train_stats(self, x, R, C, gamma, beta, rmean, rvar, momentum, eps, mean, rstd, scale, shift, ws, n_valid, stream):
        Rv = n_valid if n_valid > 0 else R
        xd = x.detach().double().view(R, C)
        m = xd.sum(0) / Rv
        var = ((xd * xd).sum(0) / Rv - m * m).clamp(min=0)
        mean.copy_(m.float())
        rstd.copy_((1.0 / torch.sqrt(var + eps)).float())
        g = gamma.detach() if gamma is not None else torch.ones(C)
        b = beta.detach() if beta is not None else torch.zeros(C)
        scale.copy_(g * rstd)
        shift.copy_(b - mean * g * rstd)
        if rmean is not None:
            rmean.mul_(1 - momentum).add_(momentum * m.float())
            rvar.mul_(1 - momentum).add_(momentum * (var * Rv / max(Rv - 1, 1)).float())

    def bn_eval_affine(self, gamma, beta, rmean, rvar, eps, scale, shift, C, stream):
        rs = torch.rsqrt(rvar + eps)
        g = gamma.detach() if gamma is not None else torch.ones(C)
        b = beta.detach() if beta is not None else torch.zeros(C)
        scale.copy_(g * rs)
        shift.copy_(b - rmean * g * rs)

    @staticmethod
    def _bn_sums(dy, x, mean, rstd, R, C):
        dyd, xd = dy.detach().double().view(R, C), x.detach().double().view(R, C)
        xh = (xd - mean.double()) * rstd.double()
        return dyd.sum(0), (dyd * xh).sum(0), xh

    def bn_bwd(self, dy, x, mean, rstd, gamma, dx, dgamma, dbeta, R, C, ws, n_valid, stream):
        Rv = n_valid if n_valid > 0 else R
        s1, s2, xh = self._bn_sums(dy, x, mean, rstd, R, C)
        g = gamma.detach().double() if gamma is not None else torch.ones(C, dtype=torch.float64)
        out = g * rstd.double() * (dy.detach().double().view(R, C) - s1 / Rv - xh * s2 / Rv)
        dx.view(R, C).copy_(out.float())
        if dgamma is not None:
            dgamma.copy_(s2.float())
        if dbeta is not None:
            dbeta.copy_(s1.float())

    def bn_bwd_coef(self, dy, x, mean, rstd, gamma, coef, dgamma, dbeta, R, C, ws, n_valid, stream):
        Rv = n_valid if n_valid > 0 else R
        s1, s2, _ = self._bn_sums(dy, x, mean, rstd, R, C)
        g = gamma.detach().double() if gamma is not None else torch.ones(C, dtype=torch.float64)
        rs, m = rstd.double(), mean.double()
        coef[:C] = (g * rs).float()
        coef[C:2 * C] = (-g * rs * rs * s2 / Rv).float()
        coef[2 * C:] = (g * rs * (m * rs * s2 - s1) / Rv).float()
        if dgamma is not None:
            dgamma.copy_(s2.float())
        if dbeta is not None:
            dbeta.copy_(s1.float())

    def log_softmax_fwd(self, x, pitch, y, R, C, stream):
        y.view(R, C).copy_(F.log_softmax(x.detach()[:R, :C].float(), -1))

    def log_softmax_bwd(self, g, y, dx, R, C, stream):
        gv, yv = g.view(R, C), y.view(R, C)
        dx.view(R, C).copy_(gv - yv.exp() * gv.sum(-1, keepdim=True))

    # ---- CNN front ------------------------------------------------------------------------------------------------------
    def conv2d_fwd(self, x, w, bias, y, N, Hi, Wi, Cin, Cout, Ho, Wo, kh, kw, sh, sw, ph, pw, stream):
        xin = x.detach().reshape(N, Hi, Wi, Cin).permute(0, 3, 1, 2)
        out = F.conv2d(xin, w.detach(), bias.detach() if bias is not None else None, stride=(sh, sw), padding=(ph, pw))
        y.copy_(out.permute(0, 2, 3, 1).reshape(N * Ho * Wo, Cout))

    def conv2d_wgrad(self, x, dy, dw, ws, N, Hi, Wi, Cin, Cout, Ho, Wo, kh, kw, sh, sw, ph, pw, stream):
        xin = x.detach().reshape(N, Hi, Wi, Cin).permute(0, 3, 1, 2)
        g = dy.detach().view(N, Ho, Wo, Cout).permute(0, 3, 1, 2)
        dw.copy_(torch.nn.grad.conv2d_weight(xin, (Cout, Cin, kh, kw), g, stride=(sh, sw), padding=(ph, pw)))

    def conv2d_dgrad(self, dy, w, dx, N, Hi, Wi, Cin, Cout, Ho, Wo, kh, kw, sh, sw, ph, pw, stream):
        g = dy.detach().view(N, Ho, Wo, Cout).permute(0, 3, 1, 2)
        d = torch.nn.grad.conv2d_input((N, Cin, Hi, Wi), w.detach(), g, stride=(sh, sw), padding=(ph, pw))
        dx.copy_(d.permute(0, 2, 3, 1))

    @staticmethod
    def _strided(t, sn, sh, sw, sc, N, Ho, Wo, C):
        return torch.as_strided(t, (N, Ho, Wo, C), (sn, sh, sw, sc))

    def affine_act(self, y, scale, shift, a, sn, sh, sw, sc, N, Ho, Wo, C, act, stream):
        v = y.detach().reshape(N, Ho, Wo, C)
        if scale is not None:
            v = v * scale + shift
        v = [torch.relu, torch.tanh, torch.sigmoid, lambda z: z][act](v)
        self._strided(a, sn, sh, sw, sc, N, Ho, Wo, C).copy_(v)

    def act_bwd_gather(self, da, a, dz, sn, sh, sw, sc, N, Ho, Wo, C, act, stream):
        d = self._strided(da, sn, sh, sw, sc, N, Ho, Wo, C)
        if act == 3:
            out = d
        else:
            av = self._strided(a, sn, sh, sw, sc, N, Ho, Wo, C)
            out = d * [(av > 0).float(), 1 - av * av, av * (1 - av)][act]
        dz.view(N, Ho, Wo, C).copy_(out)

    def maxpool2d_fwd(self, x, y, idx, N, H, W, C, kh, kw, stream):
        xin = x.detach().view(N, H, W, C).permute(0, 3, 1, 2)
        out, ind = F.max_pool2d(xin, (kh, kw), return_indices=True)
        y.copy_(out.permute(0, 2, 3, 1))
        Ho, Wo = H // kh, W // kw
        hh, ww = ind // W, ind % W                                              # absolute position of the maximum
        ho = torch.arange(Ho).view(1, 1, Ho, 1)
        wo = torch.arange(Wo).view(1, 1, 1, Wo)
        idx.copy_(((hh - ho * kh) * kw + (ww - wo * kw)).permute(0, 2, 3, 1).to(torch.uint8))

    def maxpool2d_bwd(self, dy, idx, dx, N, H, W, C, kh, kw, stream):
        Ho, Wo = H // kh, W // kw
        out = torch.zeros(N, H, W, C)
        for r in range(kh):
            for q in range(kw):
                sel = (idx == r * kw + q).float() * dy
                out[:, r:Ho * kh:kh, q:Wo * kw:kw, :] = sel
        dx.copy_(out)

    def col_sum(self, y, out, R, C, stream):
        out.copy_(y.detach().view(R, C).sum(0))

    def add_bias_rows(self, y, bias, R, C, stream):
        y.add_(bias.detach())


class _Dll(object):
    """size queries the host code makes through `_lib.lib().dll`"""

    @staticmethod
    def ctcb200_lstm_scratch_bytes(N, H):
        return 1024

    @staticmethod
    def ctcb200_conv2d_wgrad_ws_bytes(Cin, Cout, kh, kw):
        return 16

    @staticmethod
    def ctcb200_lstm_bwd_ctas(N, H, tile):
        return 64


class _FakeLib(object):
    def __init__(self, emu):
        self.emu, self.dll, self.launches = emu, _Dll(), 0

    def call(self, name, *args):
        return self.emu.call(name, *args)


@contextlib.contextmanager
def emulated():
    """Patch ctc_pytorch_b200._lib so that the model code runs on CPU tensors against the emulator."""
    from ctc_pytorch_b200 import _lib
    emu = Emu()
    fake = _FakeLib(emu)
    saved = dict(lib=_lib.lib, ptr=_lib.ptr, stream=_lib.stream, require_cuda=_lib.require_cuda)
    cuda_saved = dict(device=torch.cuda.device, current_stream=torch.cuda.current_stream)

    class _Stream(object):
        def __ne__(self, other):
            return False

        def __eq__(self, other):
            return True

    _lib.lib = lambda: fake
    _lib.ptr = lambda t: t
    _lib.stream = lambda: None
    _lib.require_cuda = lambda *a: None
    torch.cuda.device = lambda d: contextlib.nullcontext()
    torch.cuda.current_stream = lambda d=None: _Stream()
    try:
        yield emu
    finally:
        _lib.lib, _lib.ptr, _lib.stream, _lib.require_cuda = saved["lib"], saved["ptr"], saved["stream"], saved["require_cuda"]
        torch.cuda.device, torch.cuda.current_stream = cuda_saved["device"], cuda_saved["current_stream"]
