"""CUDA execution of the reference's optional CNN front (`self.conv`, timit/models/model_ctc.py:92-116,148):
a stack of LayerCNN blocks = Conv2d(bias) -> BatchNorm2d -> activation -> [MaxPool2d] -> Dropout.

Each convolution is a direct fp32 kernel (csrc/conv.cu: forward, weight gradient, data gradient from shared-memory
tiles — exact fp32 like the reference's nn.Conv2d in both precision modes); BatchNorm2d statistics are the
row-statistics kernels over the M = N*Ho*Wo rows; ReLU is fused with the BatchNorm apply. The shipped config (conf/ctc_config.yaml:32-40) uses 2-D convolutions, ReLU and no pooling;
the reference's other LayerCNN options run too — tanh / sigmoid activations (train_ctc.py:21) and nn.MaxPool2d(pool)
(model_ctc.py:53-54); Conv1d blocks raise instead of silently falling back.
"""
import torch
import torch.nn as nn

from . import _lib
from . import ops
from .model import _bn_prepare, _call, _dropout_mask, _inv_keep, _round_up


def _geometry(conv, Hi, Wi):
    kh, kw = conv.kernel_size
    sh, sw = conv.stride
    ph, pw = conv.padding
    Ho = (Hi + 2 * ph - kh) // sh + 1
    Wo = (Wi + 2 * pw - kw) // sw + 1
    return kh, kw, sh, sw, ph, pw, Ho, Wo


_ACT_CODES = {nn.ReLU: 0, nn.Tanh: 1, nn.Sigmoid: 2}   # train_ctc.py:21 supported_activate
_ACT_IDENTITY = 3


def _check_block(block):
    """Returns (activation code, pooling window or None)."""
    if not isinstance(block.conv, nn.Conv2d):
        raise RuntimeError("the B200 path implements Conv2d blocks only")
    if type(block.activation) not in _ACT_CODES:
        raise RuntimeError("the B200 path implements the reference's activations relu / tanh / sigmoid (got %r)" % (block.activation,))
    if block.conv.dilation != (1, 1) or block.conv.groups != 1:
        raise RuntimeError("the B200 path implements dense, undilated convolutions only")
    pool = None
    if block.pooling is not None:
        if not isinstance(block.pooling, nn.MaxPool2d):
            raise RuntimeError("the B200 path implements MaxPool2d pooling only")
        mp = block.pooling
        k = mp.kernel_size if isinstance(mp.kernel_size, (tuple, list)) else (mp.kernel_size, mp.kernel_size)
        st = mp.stride if isinstance(mp.stride, (tuple, list)) else (mp.stride, mp.stride)
        pd = mp.padding if isinstance(mp.padding, (tuple, list)) else (mp.padding, mp.padding)
        if tuple(st) != tuple(k) or tuple(pd) != (0, 0) or mp.dilation not in (1, (1, 1)) or mp.ceil_mode:
            raise RuntimeError("the B200 path implements nn.MaxPool2d(pool) as the reference builds it (stride = kernel, no padding)")
        pool = (int(k[0]), int(k[1]))
    return _ACT_CODES[type(block.activation)], pool


class _ConvFrontFn(torch.autograd.Function):
    """x [N, T, F] -> features [N, T', Cc*F'] in the reference's order (feature index c*F' + f)."""

    @staticmethod
    def forward(ctx, model, x, need_grad, *params):
        dev = x.device
        stream = _lib.stream
        blocks = list(model.conv.children())
        training = model.training
        N, Hi, Wi = x.shape
        act = x.contiguous()  # [N, H, W, C=1] channel-last view of the input
        Cin = 1
        saved = []
        for bi, block in enumerate(blocks):
            act_code, pool = _check_block(block)
            conv = block.conv
            if conv.in_channels != Cin:
                raise RuntimeError("conv block %d expects %d input channels, got %d" % (bi, conv.in_channels, Cin))
            kh, kw, sh, sw, ph, pw, Ho, Wo = _geometry(conv, Hi, Wi)
            Cout = conv.out_channels
            M = N * Ho * Wo
            geom = (N, Hi, Wi, Cin, Cout, Ho, Wo, kh, kw, sh, sw, ph, pw)
            y = torch.empty((M, Cout), dtype=torch.float32, device=dev)
            _call("ctcb200_conv2d_fwd", _lib.ptr(act), _lib.ptr(conv.weight), _lib.ptr(conv.bias), _lib.ptr(y), *geom, stream())
            bn = block.batch_norm
            st = _bn_prepare(bn, y, M, Cout, training) if bn is not None else None
            last = bi == len(blocks) - 1

            def _final_layout(Hx, Wx):
                if last:   # [N, H, Cout, W]: feature index c*W + w, the order the reference feeds its RNN stack
                    return torch.empty((N, Hx, Cout, Wx), dtype=torch.float32, device=dev), (Hx * Cout * Wx, Cout * Wx, 1, Wx)
                return torch.empty((N, Hx, Wx, Cout), dtype=torch.float32, device=dev), (Hx * Wx * Cout, Wx * Cout, Cout, 1)

            pool_idx = None
            if pool is None:
                a_act, a_strides = _final_layout(Ho, Wo)   # activation output directly in the block's output layout
                _call("ctcb200_affine_act", _lib.ptr(y), _lib.ptr(st.scale if st else None), _lib.ptr(st.shift if st else None),
                      _lib.ptr(a_act), a_strides[0], a_strides[1], a_strides[2], a_strides[3], N, Ho, Wo, Cout, act_code, stream())
                out, strides, Hq, Wq = a_act, a_strides, Ho, Wo
            else:
                a_act = torch.empty((N, Ho, Wo, Cout), dtype=torch.float32, device=dev)   # channel-last for the pooling
                a_strides = (Ho * Wo * Cout, Wo * Cout, Cout, 1)
                _call("ctcb200_affine_act", _lib.ptr(y), _lib.ptr(st.scale if st else None), _lib.ptr(st.shift if st else None),
                      _lib.ptr(a_act), a_strides[0], a_strides[1], a_strides[2], a_strides[3], N, Ho, Wo, Cout, act_code, stream())
                Hq, Wq = Ho // pool[0], Wo // pool[1]
                if Hq < 1 or Wq < 1:
                    raise RuntimeError("MaxPool2d%r does not fit the %dx%d activation of conv block %d" % (pool, Ho, Wo, bi))
                pooled = torch.empty((N, Hq, Wq, Cout), dtype=torch.float32, device=dev)
                pool_idx = torch.empty((N, Hq, Wq, Cout), dtype=torch.uint8, device=dev)
                _call("ctcb200_maxpool2d_fwd", _lib.ptr(a_act), _lib.ptr(pooled), _lib.ptr(pool_idx), N, Ho, Wo, Cout, pool[0],
                      pool[1], stream())
                if last:   # re-lay out for the RNN stack
                    out, strides = _final_layout(Hq, Wq)
                    _call("ctcb200_affine_act", _lib.ptr(pooled), None, None, _lib.ptr(out), strides[0], strides[1], strides[2],
                          strides[3], N, Hq, Wq, Cout, _ACT_IDENTITY, stream())
                else:
                    out, strides = pooled, (Hq * Wq * Cout, Wq * Cout, Cout, 1)
            mask = None
            p_drop = float(block.dropout.p)
            if training and p_drop > 0.0:
                mask = _dropout_mask(model, out.shape, p_drop, dev)
                if out is a_act and act_code != 0 and need_grad:
                    out = out.clone()   # tanh / sigmoid derivatives need the activation's own output, not its dropped copy
                _call("ctcb200_dropout_apply", _lib.ptr(out), _lib.ptr(mask), _inv_keep(p_drop), out.numel(), stream())
            if need_grad:
                saved.append(dict(geom=geom, Cout=Cout, M=M, x=act, y=y, st=st, a_act=a_act, a_strides=a_strides, act=act_code,
                                  pool=pool, pool_idx=pool_idx, out_strides=strides, Hq=Hq, Wq=Wq, last=last, mask=mask,
                                  p_drop=p_drop))
            act, Hi, Wi, Cin = out, Hq, Wq, Cout
        ctx.saved = saved if need_grad else None
        ctx.model = model
        ctx.param_list = params
        ctx.n_blocks = len(blocks)
        Nn, Ho, Cc, Wo = act.shape
        return act.view(Nn, Ho, Cc * Wo)

    @staticmethod
    def backward(ctx, g_out):
        saved, model = ctx.saved, ctx.model
        if saved is None:
            raise RuntimeError("backward through a forward pass that ran without gradient bookkeeping")
        dev = g_out.device
        stream = _lib.stream
        blocks = list(model.conv.children())
        grads = {}
        da = g_out.detach().to(torch.float32).contiguous()  # same strided layout as the block's output
        for bi in range(len(blocks) - 1, -1, -1):
            block, rec = blocks[bi], saved[bi]
            N, Hi, Wi, Cin, Cout, Ho, Wo, kh, kw, sh, sw, ph, pw = rec["geom"]
            M = rec["M"]
            if rec["mask"] is not None:
                _call("ctcb200_dropout_apply", _lib.ptr(da), _lib.ptr(rec["mask"]), _inv_keep(rec["p_drop"]), da.numel(),
                      stream())
            if rec["pool"] is not None:
                Hq, Wq = rec["Hq"], rec["Wq"]
                if rec["last"]:   # [N, Hq, Cout, Wq] gradient back to channel-last rows
                    os_ = rec["out_strides"]
                    dp = torch.empty((N, Hq, Wq, Cout), dtype=torch.float32, device=dev)
                    _call("ctcb200_act_bwd_gather", _lib.ptr(da), None, _lib.ptr(dp), os_[0], os_[1], os_[2], os_[3], N, Hq, Wq,
                          Cout, _ACT_IDENTITY, stream())
                else:
                    dp = da
                da = torch.empty((N, Ho, Wo, Cout), dtype=torch.float32, device=dev)
                _call("ctcb200_maxpool2d_bwd", _lib.ptr(dp), _lib.ptr(rec["pool_idx"]), _lib.ptr(da), N, Ho, Wo, Cout,
                      rec["pool"][0], rec["pool"][1], stream())
            dz = torch.empty((M, Cout), dtype=torch.float32, device=dev)
            s = rec["a_strides"]
            _call("ctcb200_act_bwd_gather", _lib.ptr(da), _lib.ptr(rec["a_act"]), _lib.ptr(dz), s[0], s[1], s[2], s[3], N, Ho,
                  Wo, Cout, rec["act"], stream())
            bn = block.batch_norm
            if bn is not None:
                dgam = torch.empty(Cout, dtype=torch.float32, device=dev)
                dbet = torch.empty(Cout, dtype=torch.float32, device=dev)
                ws = torch.empty(2 * Cout, dtype=torch.float64, device=dev)
                st = rec["st"]
                if st.batch:
                    _call("ctcb200_bn_bwd", _lib.ptr(dz), _lib.ptr(rec["y"]), _lib.ptr(st.mean), _lib.ptr(st.rstd),
                          _lib.ptr(bn.weight), _lib.ptr(dz), _lib.ptr(dgam), _lib.ptr(dbet), M, Cout, _lib.ptr(ws), 0, stream())
                else:   # frozen statistics (eval-mode fine-tuning): dx = gamma * rstd * dy
                    coef = torch.empty(3 * Cout, dtype=torch.float32, device=dev)
                    _call("ctcb200_bn_bwd_coef", _lib.ptr(dz), _lib.ptr(rec["y"]), _lib.ptr(st.mean), _lib.ptr(st.rstd),
                          _lib.ptr(bn.weight), _lib.ptr(coef), _lib.ptr(dgam), _lib.ptr(dbet), M, Cout, _lib.ptr(ws), 0, stream())
                    dz.mul_(coef[:Cout])
                grads[bn.weight], grads[bn.bias] = dgam, dbet
            conv = block.conv
            if conv.bias is not None:
                db = torch.empty(Cout, dtype=torch.float32, device=dev)
                _call("ctcb200_col_sum", _lib.ptr(dz), _lib.ptr(db), M, Cout, stream())
                grads[conv.bias] = db
            # dW[o, c, r, s] = sum_m dz[m, o] * patch(m)[r, s, c]; per-CTA partial sums reduced in a fixed order
            dw = torch.empty_like(conv.weight)
            wws = torch.empty(_lib.lib().dll.ctcb200_conv2d_wgrad_ws_bytes(Cin, Cout, kh, kw), dtype=torch.uint8, device=dev)
            _call("ctcb200_conv2d_wgrad", _lib.ptr(rec["x"]), _lib.ptr(dz), _lib.ptr(dw), _lib.ptr(wws), *rec["geom"], stream())
            grads[conv.weight] = dw
            if bi > 0:
                da = torch.empty((N, Hi, Wi, Cin), dtype=torch.float32, device=dev)
                _call("ctcb200_conv2d_dgrad", _lib.ptr(dz), _lib.ptr(conv.weight), _lib.ptr(da), *rec["geom"], stream())
        ctx.saved = None
        return (None, None, None) + tuple(grads.get(p) for p in ctx.param_list)


def conv_params(model):
    plist = []
    for block in model.conv.children():
        plist.append(block.conv.weight)
        if block.conv.bias is not None:
            plist.append(block.conv.bias)
        if block.batch_norm is not None:
            plist += [block.batch_norm.weight, block.batch_norm.bias]
    return plist


def conv_front(model, x, need_grad):
    """Features [N, T', Cc*F'] for the RNN stack (what model_ctc.py:148-156 produces before the final transpose)."""
    if x.dtype != torch.float32:
        x = x.float()
    return _ConvFrontFn.apply(model, x.contiguous(), need_grad, *conv_params(model))
