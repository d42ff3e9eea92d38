"""CPU tests of the training-loop services (ctc_pytorch_b200/train.py): the lr / rollback schedule against the oracle's
restatement of train_ctc.py:162-231 AND against the reference's own loop source exec'ed with stubs (oracle/train_live.py),
and checkpoint interchange with the unmodified reference when it is mounted."""
import random

import pytest
import torch
import torch.nn as nn

from ctc_pytorch_b200.model import CTC_Model
from ctc_pytorch_b200.train import DevLossSchedule, load_package
from oracle import ref_shim, train_ref

HAVE_REF = ref_shim.available()


def _sequence(seed, n):
    rng = random.Random(seed)
    loss, out_l, out_a = 40.0 + rng.random() * 10, [], []
    for i in range(n):
        r = rng.random()
        if r < 0.45:
            loss -= rng.random() * 2.0          # clear improvement
        elif r < 0.8:
            loss += (rng.random() - 0.5) * 0.02  # plateau inside the +-end_adjust_acc band
        else:
            loss += rng.random() * 3.0          # divergence -> immediate halving
        out_l.append(loss)
        out_a.append(rng.random())
    return out_l, out_a


@pytest.mark.parametrize("seed", range(12))
def test_schedule_matches_reference_policy(seed):
    losses, accs = _sequence(seed, 120)
    init_lr, decay, eps, epochs = 1e-3, 0.5, 0.05, 100
    want, final = train_ref.schedule_trace(losses, accs, init_lr, decay, eps, epochs)
    sched = DevLossSchedule(init_lr, decay, eps)
    got, count = [], 0
    while not sched.stop and count < epochs and count < len(losses):
        count += 1
        sched.begin_epoch()
        act = sched.update(losses[count - 1], accs[count - 1])
        got.append(dict(act, lr=sched.learning_rate))
    assert len(got) == len(want) == final["epochs"]
    for g, w in zip(got, want):
        assert g == w
    assert sched.loss_best == final["loss_best"] and sched.acc_best == final["acc_best"]
    assert any(w["rollback"] for w in want)   # the sequences do exercise the halving path


@pytest.mark.skipif(not HAVE_REF, reason="reference modules neither mounted nor staged")
@pytest.mark.parametrize("seed", range(12))
def test_schedule_pinned_to_the_reference_loop_source(seed):
    """The policy's oracle is the reference's OWN statements: oracle/train_live.py cuts the `while not stop_train:` loop out of the
    unmodified steps/train_ctc.py (lines 159-231), execs it with stubs, and both the restatement (oracle/train_ref.py) and
    DevLossSchedule must reproduce its learning rates, roll-backs and final bookkeeping epoch by epoch."""
    from oracle import train_live
    losses, accs = _sequence(seed, 120)
    losses[0] = 30.0           # below the initial loss_best = 1000 so that the counter is bound, as in any real run
    init_lr, decay, eps, epochs = 1e-3, 0.5, 0.05, 100
    live, final = train_live.reference_schedule_trace(losses, accs, init_lr, decay, eps, epochs)
    restated, rfinal = train_ref.schedule_trace(losses, accs, init_lr, decay, eps, epochs)
    sched = DevLossSchedule(init_lr, decay, eps)
    ours, count = [], 0
    while not sched.stop and count < epochs and count < len(losses):
        count += 1
        sched.begin_epoch()
        act = sched.update(losses[count - 1], accs[count - 1])
        ours.append(dict(act, lr=sched.learning_rate))
    assert len(live) == len(restated) == len(ours) == final["epochs"] == rfinal["epochs"]
    for a, b, c in zip(live, restated, ours):
        assert a["lr"] == pytest.approx(b["lr"], rel=0, abs=1e-18) and a["lr"] == pytest.approx(c["lr"], rel=0, abs=1e-18)
        assert a["rollback"] == b["rollback"] == c["rollback"]
    assert final["loss_best"] == rfinal["loss_best"] == sched.loss_best
    assert final["acc_best"] == rfinal["acc_best"] == sched.acc_best
    assert any(a["rollback"] for a in live)


def test_schedule_reproduces_the_unbound_counter_quirk():
    # first epoch inside the +-band around the initial 1000: the reference increments a variable it never assigned
    sched = DevLossSchedule(1e-3, 0.5, 0.05)
    with pytest.raises(UnboundLocalError):
        sched.update(1000.0, 0.1)
    with pytest.raises(UnboundLocalError):
        train_ref.schedule_trace([1000.0], [0.1], 1e-3, 0.5, 0.05, 5)


RNN_PARAM = {"rnn_input_size": 40, "rnn_hidden_size": 128, "rnn_layers": 2, "rnn_type": nn.LSTM, "bidirectional": True,
             "batch_norm": True}


def test_package_roundtrip_cpu():
    m = CTC_Model(rnn_param=RNN_PARAM, num_class=11, drop_out=0.1)
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    pkg = CTC_Model.save_package(m, optimizer=opt, epoch={"epoch": 3, "feature_type": "fbank", "n_feats": 40},
                                 loss_results=[3.0, 2.0, 1.5], dev_loss_results=[3.1, 2.2, 1.9], dev_cer_results=[0.1, 0.3, 0.4])
    m2, pkg2 = load_package(pkg, device="cpu")
    assert pkg2 is pkg and m2.num_class == 11 and m2.drop_out == 0.1
    for (k1, v1), (k2, v2) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2)


@pytest.mark.skipif(not HAVE_REF, reason="reference tree not mounted")
def test_checkpoint_interchange_with_reference():
    ref = ref_shim.load()
    ours = CTC_Model(rnn_param=RNN_PARAM, num_class=11, drop_out=0.1)
    theirs = ref.CTC_Model(rnn_param=RNN_PARAM, num_class=11, drop_out=0.1)
    # ours -> reference (test_ctc.py:37-60 style)
    pkg = CTC_Model.save_package(ours, epoch={"epoch": 1})
    rebuilt = ref.CTC_Model(rnn_param=pkg["rnn_param"], add_cnn=pkg["add_cnn"], cnn_param=pkg["cnn_param"],
                            num_class=pkg["num_class"], drop_out=pkg["_drop_out"])
    rebuilt.load_state_dict(pkg["state_dict"])
    # reference -> ours
    rpkg = ref.CTC_Model.save_package(theirs, epoch={"epoch": 1})
    assert set(rpkg.keys()) == set(pkg.keys())
    m2, _ = load_package(rpkg, device="cpu")
    for (k1, v1), (k2, v2) in zip(theirs.state_dict().items(), m2.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2)
