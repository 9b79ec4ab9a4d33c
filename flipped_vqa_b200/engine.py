"""Train / validation loops with the contracts of `/root/reference/engine.py:10-56` and `:59-145`
(loss-argmin branch, `:87-93,123-129`): same arguments, same returned `{meter: global_avg}` dict.

Host trimming relative to the reference (SURVEY.md §7 stage 7): the four `.item()` calls plus the
`cuda.synchronize()` per step (`engine.py:28-31,43`) become ONE device->host read of a 4-float tensor."""
from __future__ import annotations

import gc
import math
import os
import sys
from typing import Iterable

import torch

from .util import lr_sched, misc


def train_one_epoch(model: torch.nn.Module, data_loader: Iterable, optimizer: torch.optim.Optimizer, epoch: int, loss_scaler, args=None):
    model.train(True)
    metric_logger = misc.MetricLogger(delimiter="  ")
    metric_logger.add_meter("lr", misc.SmoothedValue(window_size=1, fmt="{value:.6f}"))
    header = "Epoch: [{}]".format(epoch)
    print_freq = max(int(len(data_loader) / 4), 1)
    accum_iter = args.accum_iter
    optimizer.zero_grad()
    # Long-lived objects (model, optimizer state, loader) out of the collector's reach: a generational GC pause on one rank
    # stalls every rank at the next gradient all-reduce (measured: one 94-154 ms step among 74 ms ones, bench.py step_ms_rank0)
    gc.collect()
    gc.freeze()
    for data_iter_step, data in enumerate(metric_logger.log_every(data_loader, print_freq, header)):
        if data_iter_step % accum_iter == 0:
            lr_sched.adjust_learning_rate(optimizer, data_iter_step / len(data_loader) + epoch, args)
        update = (data_iter_step + 1) % accum_iter == 0
        if hasattr(model, "require_backward_grad_sync"):      # dp.DataParallel: all-reduce only on the accumulation boundary
            model.require_backward_grad_sync = update
        if isinstance(data, tuple):                  # dataloader.PlannedLoader: (batch dict, device-resident plan)
            data, plan = data
            vqa_loss, vaq_loss, qav_loss = model.forward_plan(plan)
        else:
            vqa_loss, vaq_loss, qav_loss = model(data)
        loss = vqa_loss + vaq_loss + qav_loss
        stacked = torch.stack([loss.detach().float().reshape(()), vqa_loss.detach().float().reshape(()),
                               vaq_loss.detach().float().reshape(()), qav_loss.detach().float().reshape(())])
        vals = []

        def read_and_check():
            # one D2H read for all four logged values (the step's only host sync). A non-finite loss must never reach the
            # trainables or the AdamW state (`engine.py:28-35` exits before backward): the read happens before the optimizer
            # is touched - with our scaler AFTER backward has been enqueued, so the host waits for the forward pass while
            # the GPU is already running backward instead of idling until the host has launched it
            vals.extend(stacked.tolist())
            if not math.isfinite(vals[0]):
                print("Loss is {}, stopping training".format(vals[0]))
                sys.exit(1)

        if getattr(loss_scaler, "accepts_before_step", False):
            loss_scaler(loss / accum_iter, optimizer, parameters=model.parameters(), update_grad=update, before_step=read_and_check)
        else:                                        # any other scaler with the reference's signature: check first, like the reference
            read_and_check()
            loss_scaler(loss / accum_iter, optimizer, parameters=model.parameters(), update_grad=update)
        loss_value = vals[0]
        if update:
            optimizer.zero_grad()
        metric_logger.update(loss=loss_value, vqa_loss=vals[1], vaq_loss=vals[2], qav_loss=vals[3])
        metric_logger.update(lr=optimizer.param_groups[0]["lr"])
        if getattr(args, "debug", False):
            break
    gc.unfreeze()                                    # back into the collector's generations until the next epoch freezes again
    metric_logger.synchronize_between_processes()
    print("Averaged stats:", metric_logger)
    return {k: meter.global_avg for k, meter in metric_logger.meters.items()}


def val_one_epoch(model: torch.nn.Module, data_loader: Iterable, optimizer: torch.optim.Optimizer, epoch: int, args=None):
    """Loss-based multiple-choice accuracy (`engine.py:87-93,123-129`): the model returns per-token option
    losses; prediction = argmin over options of sum / count(loss != 0), done by one kernel."""
    model.eval()
    metric_logger = misc.MetricLogger(delimiter="  ")
    metric_logger.add_meter("lr", misc.SmoothedValue(window_size=1, fmt="{value:.6f}"))
    header = "Epoch: [{}]".format(epoch)
    print_freq = max(int(len(data_loader) / 4), 1)
    inner = model.module if hasattr(model, "module") else model
    generation = bool(getattr(args, "is_generation_task", False))
    if generation and not hasattr(inner, "generate_answers"):
        raise NotImplementedError("is_generation_task: this model has no generation evaluator (`engine.py:78-85,99-121`)")
    for data_iter_step, data in enumerate(metric_logger.log_every(data_loader, print_freq, header)):
        plan = None
        if isinstance(data, tuple):                  # dataloader.PlannedLoader(..., inference=True)
            data, plan = data
        answer = data["answer"]
        bsz = answer.shape[0]
        if generation:                               # `engine.py:78-85,99-121`: greedy generation + nearest-option matching
            with torch.no_grad():
                most_similar, extracted_answers = inner.generate_answers(data)
            if getattr(args, "output_dir", None):
                out_dir = os.path.join(args.output_dir, "extracted_answers")
                os.makedirs(out_dir, exist_ok=True)
                misc.save_result(extracted_answers, out_dir, "extracted_answers_epoch%d" % epoch)
            if getattr(args, "dataset", None) == "musicavqa":             # exact-prefix match against the first option's text
                hits = torch.tensor([int(g["generated_answer"].startswith(c["options"][0])) for c, g in zip(data["text"], extracted_answers)],
                                    dtype=torch.int32)
            else:
                hits = (answer.cpu() == most_similar.cpu())
            acc = hits.sum().item() / bsz if bsz > 0 else 0
            misc.log_qtype(data, hits, metric_logger, args)
            metric_logger.update(lr=optimizer.param_groups[0]["lr"] if optimizer is not None else 0.0)
            metric_logger.update(n=bsz, acc=acc)
            if getattr(args, "debug", False):
                break
            continue
        with torch.no_grad():
            if plan is not None:
                individual_losses = inner.inference_plan(plan)
            elif hasattr(inner, "inference"):        # the loss-based scorer even if the MODEL was built with is_generation_task set
                individual_losses = inner.inference(data)
            else:
                individual_losses = model(data, inference=True)
            prediction = inner.predict_options(individual_losses)
        eval_exact_match = (answer.to(prediction.device) == prediction).cpu()
        acc = eval_exact_match.sum().item() / bsz
        misc.log_qtype(data, eval_exact_match, metric_logger, args)             # `engine.py:127`: val_<qtype> meters
        metric_logger.update(lr=optimizer.param_groups[0]["lr"] if optimizer is not None else 0.0)
        metric_logger.update(n=bsz, acc=acc)
        if getattr(args, "debug", False):
            break
    metric_logger.synchronize_between_processes()
    print("Averaged stats:", metric_logger)
    return {k: meter.global_avg for k, meter in metric_logger.meters.items()}
