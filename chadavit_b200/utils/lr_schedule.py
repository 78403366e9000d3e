"""Learning-rate schedule of the reference's pre-training runs, as a pure function of the step.

The reference wraps its optimizer in ``LinearWarmupCosineAnnealingLR`` (src/utils/lr_scheduler.py:14-150, built by
``BaseMethod.configure_optimizers``, src/methods/base.py:441-470, stepped once per optimizer step when
``scheduler.interval == "step"``).  The fused engine step (``DINO.fused_train_step(batch, lr=...)``) takes the learning rate
as an argument — it ends up in device memory next to tau for CUDA-graph replay — so the schedule is needed in closed form:

    lr = warmup_cosine_lr(step, base_lr=cfg.optimizer.lr, warmup_steps=warmup_epochs * steps_per_epoch,
                          max_steps=max_epochs * steps_per_epoch, warmup_start_lr=3e-5, eta_min=0.0)
    loss = model.fused_train_step(batch, lr=lr)

``step`` counts completed scheduler steps (the reference's ``last_epoch``): the value returned for step s is the learning rate
the optimizer uses for its (s+1)-th update.
"""
from __future__ import annotations

import math


def warmup_cosine_lr(step: int, *, base_lr: float, warmup_steps: float, max_steps: float, warmup_start_lr: float = 0.0,
                     eta_min: float = 0.0) -> float:
    """Linear warm-up from ``warmup_start_lr`` to ``base_lr`` over ``warmup_steps`` steps, then half a cosine down to ``eta_min`` at
    ``max_steps`` (closed form of lr_scheduler.py:76-150; the chainable form the reference steps through is equal to it)."""
    if step < warmup_steps:
        if warmup_steps <= 1:
            return warmup_start_lr
        return warmup_start_lr + step * (base_lr - warmup_start_lr) / (warmup_steps - 1)
    return eta_min + 0.5 * (base_lr - eta_min) * (1.0 + math.cos(math.pi * (step - warmup_steps) / (max_steps - warmup_steps)))
