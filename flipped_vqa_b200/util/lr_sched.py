"""Per-iteration LR schedule with the semantics of `/root/reference/util/lr_sched.py:9-21`:
linear warm-up over `warmup_epochs`, then half-cosine decay to `min_lr` at `epochs`; `epoch` is the
fractional epoch (`engine.py:22-23`); param groups may carry an `lr_scale`."""
import math


def adjust_learning_rate(optimizer, epoch, args):
    if epoch < args.warmup_epochs:
        lr = args.lr * epoch / args.warmup_epochs
    else:
        progress = (epoch - args.warmup_epochs) / (args.epochs - args.warmup_epochs)
        lr = args.min_lr + (args.lr - args.min_lr) * 0.5 * (1.0 + math.cos(math.pi * progress))
    for group in optimizer.param_groups:
        group["lr"] = lr * group["lr_scale"] if "lr_scale" in group else lr
    return lr
