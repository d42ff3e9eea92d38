"""CPU suite: the host-side orchestration of the acoustic model (ctc_pytorch_b200/model.py, cnn.py, packed.py) driven against a
torch-CPU emulation of the C ABI (tests/emu_lib.py: same operand layouts, bf16 / split-bf16 arithmetic contract) and compared
with the oracle. What this pins without a GPU: every buffer reaches the right entry point in the right layout and alignment in
every mode — bf16 / x3 operands, dropout with shared masks, unidirectional layers, frozen BatchNorm, the packed (863) path, the
CNN front with and without pooling — and the arithmetic contract itself (three bf16 products reach fp32-grade gradients).
The CUDA kernels are checked against the same oracles by the `-m gpu` tests."""
import pytest
import torch
import torch.nn as nn

from ctc_pytorch_b200 import synth
from ctc_pytorch_b200.model import CTC_Model
from oracle import model_ref, packed_ref
from tests.emu_lib import emulated


def relnorm(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _loss(out, tg, il, tl, N):
    return nn.CTCLoss(reduction="sum")(out, tg, il, tl) / N


def _run_pair(m, ref, x, tg, il, tl, N, input_grad=False):
    xm = x.clone().requires_grad_(input_grad)
    out = m(xm)
    loss = _loss(out, tg, il, tl, N)
    loss.backward()
    xr = x.clone().requires_grad_(input_grad)
    rout = ref(xr)
    rloss = _loss(rout, tg, il, tl, N)
    rloss.backward()
    rp = dict(ref.named_parameters())
    worst = max(relnorm(p.grad, rp[k].grad) for k, p in m.named_parameters())
    gx = relnorm(xm.grad, xr.grad) if input_grad else 0.0
    return abs(loss.item() - rloss.item()) / abs(rloss.item()), worst, gx, (out.detach() - rout.detach()).abs().max().item()


@pytest.mark.parametrize("precision,bn,H,L", [("x3", True, 128, 2), ("x3", False, 256, 2), ("bf16", True, 128, 3)])
def test_padded_model_orchestration(precision, bn, H, L):
    T, N, F, C = 14, 5, 40, 11
    torch.manual_seed(1)
    rnn_param = {"rnn_input_size": F, "rnn_hidden_size": H, "rnn_layers": L, "rnn_type": nn.LSTM, "bidirectional": True,
                 "batch_norm": bn}
    m = CTC_Model(rnn_param=rnn_param, num_class=C, drop_out=0.0)
    ref = model_ref.RefAcousticModel(F, H, L, C, batch_norm=bn)
    ref.load_state_dict(m.state_dict())
    m.precision, m.overlap_wgrad = precision, False
    x, frac, tg, tl = synth.synthetic_batch(T, N, F, C, 4, 3)
    il = (frac * T).long()
    m.train(); ref.train()
    with emulated():
        e_loss, e_grad, e_x, e_out = _run_pair(m, ref, x, tg, il, tl, N, input_grad=True)
    if precision == "x3":
        assert e_loss < 1e-5 and e_grad < 2e-4 and e_x < 2e-4 and e_out < 1e-4, (e_loss, e_grad, e_x, e_out)
    else:
        assert e_loss < 2e-3 and e_grad < 3e-2 and e_out < 5e-2, (e_loss, e_grad, e_out)
        assert e_grad > 1e-4      # the emulator really rounds operands to bf16 (otherwise this test would prove nothing)
    for k, b in m.named_buffers():
        if "running" in k:
            assert relnorm(b, dict(ref.named_buffers())[k]) < (1e-4 if precision == "x3" else 2e-2), k


def test_dropout_and_unidirectional_orchestration():
    T, N, F, H, L, C, p = 12, 4, 40, 128, 3, 9, 0.25
    for bidir in (True, False):
        torch.manual_seed(2)
        rnn_param = {"rnn_input_size": F, "rnn_hidden_size": H, "rnn_layers": L, "rnn_type": nn.LSTM, "bidirectional": bidir,
                     "batch_norm": True}
        m = CTC_Model(rnn_param=rnn_param, num_class=C, drop_out=p)
        ref = model_ref.RefAcousticModel(F, H, L, C, batch_norm=True, dropout=p, bidirectional=bidir)
        assert list(m.state_dict().keys()) == list(ref.state_dict().keys())
        ref.load_state_dict(m.state_dict())
        m.precision, m.overlap_wgrad = "x3", False
        masks = []
        gen = torch.Generator().manual_seed(5)

        def source(shape, p_, dev):
            k = (torch.rand(shape, generator=gen) >= p_).to(torch.uint8)
            masks.append(k)
            return k
        m.mask_source = source
        x, frac, tg, tl = synth.synthetic_batch(T, N, F, C, 3, 7)
        il = (frac * T).long()
        m.train(); ref.train()
        with emulated():
            out = m(x)
            loss = _loss(out, tg, il, tl, N)
            loss.backward()
        assert len(masks) == L
        for blk, k in zip(ref.rnns.children(), masks):
            blk.fixed_mask = k
        rloss = _loss(ref(x), tg, il, tl, N)
        rloss.backward()
        rp = dict(ref.named_parameters())
        worst = max(relnorm(pp.grad, rp[k].grad) for k, pp in m.named_parameters())
        assert abs(loss.item() - rloss.item()) < 1e-5 * abs(rloss.item()) and worst < 2e-4, (bidir, worst)


def test_eval_mode_backward_orchestration():
    T, N, F, H, L, C = 10, 3, 40, 128, 2, 7
    torch.manual_seed(3)
    rnn_param = {"rnn_input_size": F, "rnn_hidden_size": H, "rnn_layers": L, "rnn_type": nn.LSTM, "bidirectional": True,
                 "batch_norm": True}
    m = CTC_Model(rnn_param=rnn_param, num_class=C, drop_out=0.2)
    for mod in m.modules():
        if isinstance(mod, nn.BatchNorm1d):
            mod.running_mean.normal_(0, 0.1)
            mod.running_var.uniform_(0.5, 1.5)
    ref = model_ref.RefAcousticModel(F, H, L, C, batch_norm=True, dropout=0.2)
    ref.load_state_dict(m.state_dict())
    m.precision, m.overlap_wgrad = "x3", False
    m.eval(); ref.eval()
    x = torch.randn(N, T, F)
    with emulated():
        xm = x.clone().requires_grad_(True)
        m(xm)[:, :, 1].sum().backward()
    xr = x.clone().requires_grad_(True)
    ref(xr)[:, :, 1].sum().backward()
    assert relnorm(xm.grad, xr.grad) < 2e-4
    rp = dict(ref.named_parameters())
    for k, p in m.named_parameters():
        assert relnorm(p.grad, rp[k].grad) < 2e-4, k


@pytest.mark.parametrize("pool", [False, True])
def test_cnn_front_orchestration(pool):
    T, N, F, H, C = 24, 2, 40, 128, 9
    layers = [[(1, 8), (3, 3), (1, 2), (1, 1), (2, 1) if pool else None], [(8, 8), (3, 3), (2, 2) if not pool else (1, 1), (1, 1), None]]
    torch.manual_seed(4)
    rnn_param = {"rnn_input_size": F, "rnn_hidden_size": H, "rnn_layers": 1, "rnn_type": nn.LSTM, "bidirectional": True,
                 "batch_norm": True}
    cnn_param = {"batch_norm": True, "activate_function": nn.ReLU, "layer": layers}
    m = CTC_Model(add_cnn=True, cnn_param=cnn_param, rnn_param=rnn_param, num_class=C, drop_out=0.0)
    ref = model_ref.RefAcousticModel(F, H, 1, C, batch_norm=True, cnn_layers=layers, cnn_batch_norm=True)
    assert list(m.state_dict().keys()) == list(ref.state_dict().keys())
    ref.load_state_dict(m.state_dict())
    m.precision, m.overlap_wgrad = "x3", False
    x, frac, tg, tl = synth.synthetic_batch(T, N, F, C, 3, 5)
    m.train(); ref.train()
    with emulated():
        out = m(x)
        il = (frac * out.shape[0]).long()
        loss = _loss(out, tg, il, tl, N)
        loss.backward()
    rout = ref(x)
    assert rout.shape == out.shape
    rloss = _loss(rout, tg, il, tl, N)
    rloss.backward()
    assert abs(loss.item() - rloss.item()) < 1e-5 * abs(rloss.item())
    rp = dict(ref.named_parameters())
    for k, p in m.named_parameters():
        if k.endswith("conv.bias"):
            assert p.grad.abs().max().item() < 1e-4        # a bias in front of BatchNorm: exactly-zero gradient up to round-off
            continue
        assert relnorm(p.grad, rp[k].grad) < 5e-4, (k, relnorm(p.grad, rp[k].grad))


@pytest.mark.parametrize("precision", ["x3", "bf16"])
def test_packed_model_orchestration(precision):
    """The alignment design of the packed (863) path end to end: CTC_RNN on a ragged batch against the per-utterance oracle."""
    from ctc_pytorch_b200.packed import CTC_RNN
    T, N, F_, H, L, C = 18, 6, 40, 128, 3, 10
    torch.manual_seed(6)
    m = CTC_RNN(rnn_input_size=F_, rnn_hidden_size=H, rnn_layers=L, rnn_type=nn.LSTM, bidirectional=True, batch_norm=True,
                num_class=C, drop_out=0.0)
    ref = packed_ref.RefPackedModel(F_, H, L, True, C)
    assert list(m.state_dict().keys()) == list(ref.state_dict().keys())
    ref.load_state_dict(m.state_dict())
    m.precision = precision
    x, lens, targets, tsz = packed_ref.synthetic_packed_batch(T, N, F_, C, 3, 8)
    m.train(); ref.train()
    with emulated():
        act = m(nn.utils.rnn.pack_padded_sequence(x, lens))
        loss = packed_ref.warp_ctc_loss(act, targets, lens, tsz)
        loss.backward()
    ract = ref(x, lens)
    rloss = packed_ref.warp_ctc_loss(ract, targets, lens, tsz)
    rloss.backward()
    for n, Ln in enumerate(lens):
        assert act[Ln:, n].abs().sum().item() == 0.0            # padded frames are zero vectors
    rp = dict(ref.named_parameters())
    worst = max(relnorm(p.grad, rp[k].grad) for k, p in m.named_parameters())
    tol = (1e-5, 2e-4, 1e-4) if precision == "x3" else (2e-3, 3e-2, 5e-2)
    assert abs(loss.item() - rloss.item()) < tol[0] * abs(rloss.item()), (loss.item(), rloss.item())
    assert worst < tol[1], worst
    assert (act.detach() - ract.detach()).abs().max().item() < tol[2]
    rb = dict(ref.named_buffers())
    for k, b in m.named_buffers():
        if "running" in k:
            assert relnorm(b, rb[k]) < (1e-4 if precision == "x3" else 2e-2), k
    m.eval(); ref.eval()
    with emulated(), torch.no_grad():
        logp = m(x, lens)
    assert (logp - ref(x, lens)).abs().max().item() < tol[2]


@pytest.mark.parametrize("rnn_type,nonlin", [(nn.GRU, None), (nn.RNN, "tanh"), (nn.RNN, "relu")])
def test_gru_and_rnn_cells_orchestration(rnn_type, nonlin):
    """The reference's other recurrent cells (train_ctc.py:20 supported_rnn): same four-slot operand layout, cell-specific
    element phases; GRU's second set of gate gradients (input-side vs recurrent-side n slot) reaches the right GEMMs."""
    T, N, F, H, L, C = 11, 4, 40, 128, 2, 8
    torch.manual_seed(9)
    rnn_param = {"rnn_input_size": F, "rnn_hidden_size": H, "rnn_layers": L, "rnn_type": rnn_type, "bidirectional": True,
                 "batch_norm": True}
    m = CTC_Model(rnn_param=rnn_param, num_class=C, drop_out=0.0)
    ref = model_ref.RefAcousticModel(F, H, L, C, batch_norm=True, rnn_type=rnn_type)
    if nonlin == "relu":   # not reachable through the reference's config (nn.RNN defaults to tanh); swap the modules in
        for model_ in (m, ref):
            for blk in model_.rnns.children():
                old = blk.rnn
                blk.rnn = nn.RNN(input_size=old.input_size, hidden_size=old.hidden_size, bidirectional=True, bias=False,
                                 nonlinearity="relu")
    assert list(m.state_dict().keys()) == list(ref.state_dict().keys())
    ref.load_state_dict(m.state_dict())
    m.precision, m.overlap_wgrad = "x3", False
    x, frac, tg, tl = synth.synthetic_batch(T, N, F, C, 3, 4)
    il = (frac * T).long()
    m.train(); ref.train()
    with emulated():
        e_loss, e_grad, e_x, e_out = _run_pair(m, ref, x, tg, il, tl, N, input_grad=True)
    assert e_loss < 1e-5 and e_grad < 3e-4 and e_x < 3e-4 and e_out < 1e-4, (e_loss, e_grad, e_x, e_out)
