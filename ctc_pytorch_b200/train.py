"""Training-loop services around the step: the epoch loop, the dev-loss driven learning-rate
halving / rollback schedule with in-RAM best-state snapshots, and checkpoint interchange with the reference's
`save_package` format including a resume path.

Mirrors timit/steps/train_ctc.py: `run_epoch` (26-69) and the `while not stop_train` loop of `main` (173-249). The
arithmetic of a step is unchanged (model -> CTCLoss(sum)/batch -> backward -> optimizer.step); what changes is where the
bookkeeping lives: loss and edit-distance terms are accumulated on the device (arg-max, collapse and Levenshtein run as
kernels) and read back once per `print_every` steps and at the end of the epoch instead of two `.item()` calls and a
NumPy round trip per step (train_ctc.py:49-52).
"""
import copy
import os
import time

import torch

from . import ops
from .model import CTC_Model


class DevLossSchedule(object):
    """The reference's learning-rate policy as a state machine (train_ctc.py:162-171, 181-185, 197-231).

    After every epoch call `update(dev_loss, acc)`; it returns a dict of actions for the caller:
      snapshot   keep a copy of (model, optimizer) as the roll-back state        (train_ctc.py:201-202, 207-208)
      best       keep a copy as the best-accuracy state that is finally saved    (train_ctc.py:213-215)
      rollback   reload the roll-back state now                                  (train_ctc.py:227-228)
      stop       stop training                                                   (train_ctc.py:230-231)
    and `begin_epoch()` returns the factor to multiply every param group's lr with (1.0 or `decay`).
    """

    def __init__(self, init_lr, decay, end_adjust_acc, max_adjust=8, patience=10):
        self.learning_rate = init_lr
        self.decay = decay
        self.end_adjust_acc = end_adjust_acc
        self.max_adjust = max_adjust
        self.patience = patience
        self.loss_best = 1000
        self.loss_best_true = 1000
        self.adjust_rate_flag = False
        self.adjust_rate_count = None   # the reference leaves it unbound until the first improving epoch
        self.adjust_time = 0
        self.acc_best = 0
        self.stop = False

    def begin_epoch(self):
        if self.adjust_rate_flag:
            self.learning_rate *= self.decay
            self.adjust_rate_flag = False
            return self.decay
        return 1.0

    def update(self, dev_loss, acc):
        act = {"snapshot": False, "best": False, "rollback": False, "stop": False}
        if dev_loss < (self.loss_best - self.end_adjust_acc):
            self.loss_best = dev_loss
            self.loss_best_true = dev_loss
            self.adjust_rate_count = 0
            act["snapshot"] = True
        elif dev_loss < self.loss_best + self.end_adjust_acc:
            if self.adjust_rate_count is None:
                raise UnboundLocalError("local variable 'adjust_rate_count' referenced before assignment")
            self.adjust_rate_count += 1
            if dev_loss < self.loss_best and dev_loss < self.loss_best_true:
                self.loss_best_true = dev_loss
                act["snapshot"] = True
        else:
            self.adjust_rate_count = self.patience
        if acc > self.acc_best:
            self.acc_best = acc
            act["best"] = True
        if self.adjust_rate_count == self.patience:
            self.adjust_rate_flag = True
            self.adjust_time += 1
            self.adjust_rate_count = 0
            if self.loss_best > self.loss_best_true:
                self.loss_best = self.loss_best_true
            act["rollback"] = True
        if self.adjust_time == self.max_adjust:
            self.stop = True
            act["stop"] = True
        return act


def _step_terms(model, out, input_sizes, targets, target_sizes):
    """(errors, tokens) of one batch as device scalars (no synchronisation)."""
    _, labels, lens = ops.greedy_decode(out, input_sizes, blank=0)
    tsz = target_sizes.to(device=out.device, dtype=torch.int64)
    dist = ops.edit_distance(labels, lens, targets.to(out.device), tsz)
    return dist.sum().to(torch.float64), tsz.sum().to(torch.float64)


def run_epoch(epoch_id, model, data_iter, loss_fn, device, optimizer=None, print_every=20, is_training=True, log=print):
    """One pass over `data_iter` (batches `(inputs, input_sizes, targets, target_sizes, utt_list)` as produced by
    create_input, data_loader.py:119-140). Returns `(1 - errors/tokens, average loss)` like train_ctc.py:69."""
    if is_training:
        model.train()
    else:
        model.eval()
    acc = torch.zeros(4, dtype=torch.float64, device=device)   # total loss, window loss, errors, tokens
    steps = 0
    for i, data in enumerate(data_iter):
        inputs, input_sizes, targets, target_sizes = data[0], data[1], data[2], data[3]
        inputs = inputs.to(device, non_blocking=True)
        input_sizes = input_sizes.to(device, non_blocking=True)
        targets = targets.to(device, non_blocking=True)
        target_sizes = target_sizes.to(device, non_blocking=True)
        with torch.set_grad_enabled(is_training):
            out = model(inputs)
            out_len, batch_size, _ = out.size()
            in_len = (input_sizes * out_len).long()
            loss = loss_fn(out, targets, in_len, target_sizes) / batch_size
        errs, toks = _step_terms(model, out.detach(), in_len, targets, target_sizes)
        ld = loss.detach().to(torch.float64)
        acc += torch.stack([ld, ld, errs, toks])
        steps = i + 1
        if steps % print_every == 0 and is_training:
            tot, cur, e, t = acc.tolist()   # the only synchronisation inside the epoch
            log("Epoch = %d, step = %d, cur_loss = %.4f, total_loss = %.4f, total_wer = %.4f" % (
                epoch_id, steps, cur / print_every, tot / steps, e / t))
            acc[1] = 0.0
        if is_training:
            optimizer.zero_grad()
            loss.backward()
            optimizer.step()
    tot, _, e, t = acc.tolist()
    average_loss = tot / max(steps, 1)
    log("Epoch %d %s done, total_loss: %.4f, total_wer: %.4f" % (epoch_id, "Train" if is_training else "Valid", average_loss,
                                                                  e / max(t, 1.0)))
    return 1 - e / max(t, 1.0), average_loss


def load_package(path_or_package, device="cuda", with_optimizer=None):
    """Rebuild a CTC_Model from a checkpoint written by either side's `save_package` (model_ctc.py:209-229; the reference
    reads it back at test_ctc.py:37-60). Returns (model, package); loads `optim_dict` into `with_optimizer(model)` if given."""
    package = path_or_package
    if not isinstance(package, dict):
        package = torch.load(path_or_package, map_location="cpu", weights_only=False)
    model = CTC_Model(rnn_param=package["rnn_param"], add_cnn=package["add_cnn"], cnn_param=package["cnn_param"],
                      num_class=package["num_class"], drop_out=package["_drop_out"])
    model.load_state_dict(package["state_dict"])
    model.to(device)
    optimizer = None
    if with_optimizer is not None:
        optimizer = with_optimizer(model)
        if "optim_dict" in package:
            optimizer.load_state_dict(package["optim_dict"])
    return (model, package) if optimizer is None else (model, package, optimizer)


def fit(model, train_loader, dev_loader, loss_fn, optimizer, device, init_lr, decay, end_adjust_acc, num_epoches, params=None,
        print_every=20, checkpoint_path=None, resume=None, log=print):
    """The reference's training driver (train_ctc.py:173-249) without Visdom: epochs of run_epoch(train) + run_epoch(dev),
    learning-rate halving / rollback by DevLossSchedule, best-accuracy state restored and saved at the end.
    `resume`: a package (or path) from an earlier `fit`; restores weights, optimizer state, epoch count and histories."""
    sched = DevLossSchedule(init_lr, decay, end_adjust_acc)
    loss_results, dev_loss_results, dev_cer_results = [], [], []
    count = 0
    if resume is not None:
        package = resume if isinstance(resume, dict) else torch.load(resume, map_location="cpu", weights_only=False)
        model.load_state_dict(package["state_dict"])
        if "optim_dict" in package:
            optimizer.load_state_dict(package["optim_dict"])
        loss_results = list(package.get("loss_results") or [])
        dev_loss_results = list(package.get("dev_loss_results") or [])
        dev_cer_results = list(package.get("dev_cer_results") or [])
        ep = package.get("epoch")
        count = int(ep["epoch"]) if isinstance(ep, dict) and "epoch" in ep else len(loss_results)
        sched.learning_rate = optimizer.param_groups[0]["lr"]
    model_state = op_state = best_model_state = best_op_state = None
    start = time.time()
    while not sched.stop:
        if count >= num_epoches:
            break
        count += 1
        factor = sched.begin_epoch()
        if factor != 1.0:
            for group in optimizer.param_groups:
                group["lr"] *= factor
        log("Start training epoch: %d, learning_rate: %.5f" % (count, sched.learning_rate))
        _, loss = run_epoch(count, model, train_loader, loss_fn, device, optimizer=optimizer, print_every=print_every,
                            is_training=True, log=log)
        loss_results.append(loss)
        acc, dev_loss = run_epoch(count, model, dev_loader, loss_fn, device, optimizer=None, print_every=print_every,
                                  is_training=False, log=log)
        log("loss on dev set is %.4f" % dev_loss)
        dev_loss_results.append(dev_loss)
        dev_cer_results.append(acc)
        act = sched.update(dev_loss, acc)
        if act["snapshot"]:
            model_state = copy.deepcopy(model.state_dict())
            op_state = copy.deepcopy(optimizer.state_dict())
        if act["best"]:
            best_model_state = copy.deepcopy(model.state_dict())
            best_op_state = copy.deepcopy(optimizer.state_dict())
        if act["rollback"]:
            model.load_state_dict(model_state)
            optimizer.load_state_dict(op_state)
        log("epoch %d done, cv acc is: %.4f, time_used: %.4f minutes" % (count, acc, (time.time() - start) / 60))
    log("End training, best dev loss is: %.4f, acc is: %.4f" % (sched.loss_best, sched.acc_best))
    if best_model_state is not None:
        model.load_state_dict(best_model_state)
        optimizer.load_state_dict(best_op_state)
    params = dict(params or {})
    params["epoch"] = count
    package = CTC_Model.save_package(model, optimizer=optimizer, epoch=params, loss_results=loss_results,
                                     dev_loss_results=dev_loss_results, dev_cer_results=dev_cer_results)
    if checkpoint_path is not None:
        os.makedirs(os.path.dirname(os.path.abspath(checkpoint_path)), exist_ok=True)
        torch.save(package, checkpoint_path)
    return package, sched
