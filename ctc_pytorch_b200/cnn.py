"""CUDA execution of the reference's optional CNN front (`self.conv`, timit/models/model_ctc.py:92-116,148):
a stack of LayerCNN blocks = Conv2d(bias) -> BatchNorm2d -> activation -> [MaxPool2d] -> Dropout.

Each convolution is a direct fp32 kernel (csrc/conv.cu: forward, weight gradient, data gradient from shared-memory
tiles — exact fp32 like the reference's nn.Conv2d in both precision modes); BatchNorm2d statistics are the
row-statistics kernels over the M = N*Ho*Wo rows; ReLU is fused with the BatchNorm apply. Supported on the
CUDA path: 2-D convolutions, ReLU activation, no pooling (the shipped config, conf/ctc_config.yaml:32-40);
anything else raises instead of silently falling back.
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


def _check_block(block):
    if not isinstance(block.conv, nn.Conv2d):
        raise RuntimeError("the B200 path implements Conv2d blocks only")
    if not isinstance(block.activation, nn.ReLU):
        raise RuntimeError("the B200 path implements the ReLU activation only (got %r)" % (block.activation,))
    if block.pooling is not None:
        raise RuntimeError("the B200 path does not implement the optional MaxPool2d of LayerCNN")
    if block.conv.dilation != (1, 1) or block.conv.groups != 1:
        raise RuntimeError("the B200 path implements dense, undilated convolutions only")


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
            _check_block(block)
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
            if last:   # [N, Ho, Cout, Wo]: feature index c*Wo + w, the order the reference feeds its RNN stack
                out = torch.empty((N, Ho, Cout, Wo), dtype=torch.float32, device=dev)
                strides = (Ho * Cout * Wo, Cout * Wo, 1, Wo)
            else:      # channel-last for the next im2col
                out = torch.empty((N, Ho, Wo, Cout), dtype=torch.float32, device=dev)
                strides = (Ho * Wo * Cout, Wo * Cout, Cout, 1)
            _call("ctcb200_affine_relu", _lib.ptr(y), _lib.ptr(st.scale if st else None),
                  _lib.ptr(st.shift if st else None), _lib.ptr(out), strides[0], strides[1], strides[2], strides[3],
                  N, Ho, Wo, Cout, stream())
            mask = None
            p_drop = float(block.dropout.p)
            if training and p_drop > 0.0:
                mask = _dropout_mask(model, out.shape, p_drop, dev)
                _call("ctcb200_dropout_apply", _lib.ptr(out), _lib.ptr(mask), _inv_keep(p_drop), out.numel(), stream())
            if need_grad:
                saved.append(dict(geom=geom, Cout=Cout, M=M, x=act, y=y, st=st, out=out, strides=strides, mask=mask,
                                  p_drop=p_drop))
            act, Hi, Wi, Cin = out, Ho, Wo, Cout
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
            dz = torch.empty((M, Cout), dtype=torch.float32, device=dev)
            s = rec["strides"]
            _call("ctcb200_relu_bwd_gather", _lib.ptr(da), _lib.ptr(rec["out"]), _lib.ptr(dz), s[0], s[1], s[2], s[3], N, Ho,
                  Wo, Cout, stream())
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
