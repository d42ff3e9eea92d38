"""ORACLE — test infrastructure only. Pins the learning-rate / rollback policy to the reference's OWN source lines.

The policy lives inside `main` of timit/steps/train_ctc.py (the `while not stop_train:` loop, lines 159-231), behind Visdom and
the Kaldi data layer, so it cannot be imported. Here the loop's source text is cut out of the unmodified file (from
/root/reference, or its staged copy oracle/_ref), dedented and exec'ed with stubs for everything it touches: `run_epoch`
replays a scripted (accuracy, dev loss) sequence, `model` / `optimizer` record load_state_dict / state_dict calls, `viz` is inert.
What comes back is a per-epoch trace in the format of oracle/train_ref.schedule_trace, produced by the reference's statements.
"""
import copy
import os
import textwrap
import time

import numpy as np

from oracle import ref_shim


def _loop_source():
    path = os.path.join(ref_shim.REF_ROOT, "steps", "train_ctc.py")
    lines = open(path).read().split("\n")
    start = next(i for i, l in enumerate(lines) if l.strip() == "count = 0")
    end = next(i for i, l in enumerate(lines) if l.strip().startswith('print("End training'))
    return textwrap.dedent("\n".join(lines[start:end]))


class _Recorder(object):
    def __init__(self, name, log):
        self.name, self.log, self.version = name, log, 0
        self.param_groups = [{"lr": None}]

    def state_dict(self):
        self.version += 1
        return {"tag": "%s@%d" % (self.name, self.version)}

    def load_state_dict(self, sd):
        self.log.append(("load", self.name, sd["tag"]))


class _Viz(object):
    def line(self, **kw):
        return object()


def reference_schedule_trace(dev_losses, accs, init_lr, decay, end_adjust_acc, num_epoches):
    """Runs the reference's loop text; returns (trace, summary) like oracle/train_ref.schedule_trace, or raises what the
    reference raises (UnboundLocalError -> NameError under exec: `adjust_rate_count` read before any epoch improved)."""
    log, trace = [], []
    model, optimizer = _Recorder("model", log), _Recorder("optim", log)
    optimizer.param_groups[0]["lr"] = init_lr
    calls = {"n": 0}

    def run_epoch(epoch_id, model_, loader, loss_fn, device, optimizer=None, print_every=20, is_training=True):
        if is_training:
            # one trace row per epoch, opened when the training pass starts with the learning rate the loop applied
            trace.append({"lr": optimizer.param_groups[0]["lr"], "loads": len(log)})
            return 0.0, 0.0
        k = calls["n"]
        calls["n"] += 1
        return accs[k], dev_losses[k]

    class _Opts(object):
        verbose_step = 20
    ns = dict(init_lr=init_lr, decay=decay, end_adjust_acc=end_adjust_acc, num_epoches=min(num_epoches, len(dev_losses)),
              run_epoch=run_epoch, model=model, optimizer=optimizer, train_loader=None, dev_loader=None, loss_fn=None, device=None,
              opts=_Opts(), copy=copy, time=time, np=np, viz=_Viz(), viz_window=[None, None, None], viz_opts=[{}, {}, {}],
              print=lambda *a, **k: None)
    exec(compile(_loop_source(), "train_ctc.py:loop", "exec"), ns)
    # rebuild the per-epoch events from what the loop did to its stubs
    out = []
    for i, row in enumerate(trace):
        nxt = trace[i + 1]["loads"] if i + 1 < len(trace) else len(log)
        out.append({"lr": row["lr"], "rollback": nxt > row["loads"]})
    return out, {"loss_best": ns["loss_best"], "acc_best": ns["acc_best"], "epochs": ns["count"], "adjust_time": ns["adjust_time"]}
