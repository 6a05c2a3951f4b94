"""Learning-rate schedules of the hot path's callers.

Same class names and constructor arguments as ``nasrec/utils/lr_schedule.py`` so the
entry scripts keep working, restated around two pure functions of the step index:

* ``constant_warmup_scale(step_count, warmup)``   -- lr_schedule.py:33-42
* ``cosine_warmup_lr(step_in_cycle, ...)``        -- lr_schedule.py:98-120

The pure functions are what ``FusedTrainer``/``SubnetEvaluator`` use (they take ``lr`` per
step and own no torch optimizer); the classes wrap a ``torch.optim.Optimizer`` for the
drop-in path.  tests/test_search_cpu.py checks both against sequences generated from
the reference classes (tests/golden/lr_schedules.json).
"""
from __future__ import annotations

import math
from typing import List

import torch
from torch.optim.lr_scheduler import _LRScheduler


def constant_warmup_scale(step_count: int, num_warmup_steps: int) -> float:
    """Linear ramp k/W for k <= W, then 1 (lr_schedule.py:35-41)."""
    if step_count <= num_warmup_steps:
        return 1.0 - (num_warmup_steps - step_count) / num_warmup_steps
    return 1.0


def cosine_warmup_lr(step_in_cycle: int, cycle_steps: int, warmup_steps: int, max_lr: float, min_lr: float) -> float:
    """lr at position ``step_in_cycle`` of a cycle (lr_schedule.py:98-120); -1 is the
    'before the first step' position, which holds min_lr."""
    if step_in_cycle == -1:
        return min_lr
    if step_in_cycle < warmup_steps:
        return (max_lr - min_lr) * step_in_cycle / warmup_steps + min_lr
    phase = math.pi * (step_in_cycle - warmup_steps) / (cycle_steps - warmup_steps)
    return min_lr + (max_lr - min_lr) * (1 + math.cos(phase)) / 2


class CosineCursor:
    """Position bookkeeping of the warm-restart schedule without an optimizer:
    ``advance()`` is ``step()`` and ``seek(e)`` is ``step(epoch=e)`` (lr_schedule.py:122-159)."""

    def __init__(self, first_cycle_steps: int, cycle_mult: float = 1.0, max_lr: float = 0.1, min_lr: float = 0.001,
                 warmup_steps: int = 0, gamma: float = 1.0, last_epoch: int = -1):
        assert warmup_steps < first_cycle_steps
        self.first_cycle_steps = first_cycle_steps
        self.cycle_mult = cycle_mult
        self.base_max_lr = self.max_lr = max_lr
        self.min_lr = min_lr
        self.warmup_steps = warmup_steps
        self.gamma = gamma
        self.cur_cycle_steps = first_cycle_steps
        self.cycle = 0
        self.step_in_cycle = last_epoch
        self.last_epoch = last_epoch

    def _finish(self, epoch) -> float:
        self.max_lr = self.base_max_lr * (self.gamma ** self.cycle)
        self.last_epoch = math.floor(epoch)
        return self.lr()

    def lr(self) -> float:
        return cosine_warmup_lr(self.step_in_cycle, self.cur_cycle_steps, self.warmup_steps, self.max_lr, self.min_lr)

    def advance(self) -> float:
        self.step_in_cycle += 1
        if self.step_in_cycle >= self.cur_cycle_steps:           # restart, cycle length stretched by cycle_mult
            self.cycle += 1
            self.step_in_cycle -= self.cur_cycle_steps
            self.cur_cycle_steps = int((self.cur_cycle_steps - self.warmup_steps) * self.cycle_mult) + self.warmup_steps
        return self._finish(self.last_epoch + 1)

    def seek(self, epoch) -> float:
        first, mult = self.first_cycle_steps, self.cycle_mult
        if epoch < first:
            self.cur_cycle_steps, self.step_in_cycle = first, epoch
        elif mult == 1.0:
            self.cycle, self.step_in_cycle = epoch // first, epoch % first
        else:                                                    # geometric cycle lengths first * mult^n
            n = int(math.log(epoch / first * (mult - 1) + 1, mult))
            self.cycle = n
            self.step_in_cycle = epoch - int(first * (mult ** n - 1) / (mult - 1))
            self.cur_cycle_steps = first * mult ** n
        return self._finish(epoch)


class ConstantWithWarmup(_LRScheduler):
    """lr_schedule.py:21-42: constant base lr after a linear warm-up."""

    def __init__(self, optimizer, num_warmup_steps: int):
        self.num_warmup_steps = num_warmup_steps
        super().__init__(optimizer)

    def get_lr(self):
        scale = constant_warmup_scale(self._step_count, self.num_warmup_steps)
        if self._step_count <= self.num_warmup_steps:
            self.last_lr = [b * scale for b in self.base_lrs]
            return self.last_lr
        return self.base_lrs


class CosineAnnealingWarmupRestarts(_LRScheduler):
    """lr_schedule.py:47-164: cosine decay with linear warm-up and warm restarts; every
    param group starts at (and decays back to) ``min_lr``."""

    def __init__(self, optimizer: torch.optim.Optimizer, first_cycle_steps: int, cycle_mult: float = 1.0,
                 max_lr: float = 0.1, min_lr: float = 0.001, warmup_steps: int = 0, gamma: float = 1.0,
                 last_epoch: int = -1):
        self._cur = CosineCursor(first_cycle_steps, cycle_mult, max_lr, min_lr, warmup_steps, gamma, last_epoch)
        super().__init__(optimizer, last_epoch)
        self.base_lrs = []
        for group in self.optimizer.param_groups:                # lr_schedule.py:91-96
            group["lr"] = min_lr
            self.base_lrs.append(min_lr)

    # the attributes user code reads on the reference class
    first_cycle_steps = property(lambda self: self._cur.first_cycle_steps)
    cur_cycle_steps = property(lambda self: self._cur.cur_cycle_steps)
    step_in_cycle = property(lambda self: self._cur.step_in_cycle)
    cycle = property(lambda self: self._cur.cycle)
    max_lr = property(lambda self: self._cur.max_lr)
    min_lr = property(lambda self: self._cur.min_lr)
    warmup_steps = property(lambda self: self._cur.warmup_steps)

    def get_lr(self):
        c = self._cur
        return [cosine_warmup_lr(c.step_in_cycle, c.cur_cycle_steps, c.warmup_steps, c.max_lr, b)
                for b in self.base_lrs]

    def step(self, epoch=None):
        if epoch is None:
            self._cur.advance()
        else:
            self._cur.seek(epoch)
        self.last_epoch = self._cur.last_epoch
        for group, lr in zip(self.optimizer.param_groups, self.get_lr()):
            group["lr"] = lr


def finetune_lr_sequence(num_steps: int, max_lr: float, min_lr: float = 1e-8) -> List[float]:
    """The lr each optimizer step sees in the EA's one-shot fine-tune
    (eval_subnet_from_supernet.py:148-183 + train_utils.py:262-300): cosine schedule with
    warm-up = steps // 10, positioned with step(epoch=-1) before the loop and advanced
    once after every optimizer step."""
    cur = CosineCursor(num_steps, max_lr=max_lr, min_lr=min_lr, warmup_steps=num_steps // 10)
    cur.advance()                      # _LRScheduler.__init__ performs one step()
    lr = cur.seek(-1)
    out = []
    for _ in range(num_steps):
        out.append(lr)
        lr = cur.advance()
    return out
