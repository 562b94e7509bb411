"""Learning-rate schedules with warm-up, in closed form.

Spec: reference libai/scheduler/lr_scheduler.py:23-257 (six ``Warmup*LR`` factories wrapping a
OneFlow scheduler in ``WarmUpLR``) with the goldens of tests/test_scheduler.py:35-196.  Let ``S(k)``
be the wrapped schedule at absolute step ``k``, ``W = warmup_iter``, ``f = warmup_factor`` and ``η``
the group's base lr:

* linear warm-up **interpolates** from ``η·f`` to ``S(W)``: ``lr(k) = η·f + (S(W) − η·f)·k/W`` for
  ``k < W``; constant warm-up holds ``η·f``; afterwards ``lr(k) = S(k)``; ``W == 0`` → bare ``S``.

Each factory returns a :class:`ClosedFormLR` (a ``torch.optim.lr_scheduler.LRScheduler``), so
``step() / get_last_lr() / state_dict()`` behave as usual.
"""
from __future__ import annotations

import bisect
import logging
import math
from typing import Callable, List

from torch.optim.lr_scheduler import LRScheduler

logger = logging.getLogger(__name__)


class ClosedFormLR(LRScheduler):
    """``lr_i(k) = warmup(S_i)(k)`` where ``S_i(k) = base_lr_i · shape(k)`` (or a full function)."""

    def __init__(self, optimizer, schedule: Callable[[float, int], float], warmup_factor: float = 0.0,
                 warmup_iter: int = 0, warmup_method: str = "linear", last_epoch: int = -1):
        if warmup_method not in ("linear", "constant"):
            raise ValueError(f"Unknown warmup method: {warmup_method}")
        self._schedule = schedule
        self.warmup_factor, self.warmup_iter, self.warmup_method = warmup_factor, int(warmup_iter), warmup_method
        super().__init__(optimizer, last_epoch)

    def lr_at(self, base_lr: float, k: int) -> float:
        W = self.warmup_iter
        if W <= 0 or k >= W:
            return self._schedule(base_lr, k)
        start = base_lr * self.warmup_factor
        if self.warmup_method == "constant":
            return start
        end = self._schedule(base_lr, W)
        return start + (end - start) * k / W

    def get_lr(self) -> List[float]:
        return [self.lr_at(b, self.last_epoch) for b in self.base_lrs]

    def state_dict(self):
        return {k: v for k, v in self.__dict__.items() if k not in ("optimizer", "_schedule")}


def WarmupCosineLR(optimizer, max_iter: int, warmup_factor: float, warmup_iter: int, alpha: float = 0.0,
                   warmup_method: str = "linear"):
    """Cosine decay ``η·((1−α)·½(1+cos(π·k/max_iter)) + α)`` (held at ``η·α`` after ``max_iter``)."""

    def sched(eta, k):
        if k < max_iter:
            return eta * ((1.0 - alpha) * 0.5 * (1.0 + math.cos(math.pi * k / max_iter)) + alpha)
        return eta * alpha

    if warmup_iter == 0:
        logger.warning("warmup iters equals to zero, return CosineLR")
    return ClosedFormLR(optimizer, sched, warmup_factor, warmup_iter, warmup_method)


def WarmupCosineAnnealingLR(optimizer, max_iter: int, warmup_factor: float, warmup_iter: int, eta_min: float = 0.0,
                            warmup_method: str = "linear"):
    """Cosine annealing with ``T_max = max_iter`` down to ``eta_min`` (periodic like PyTorch's)."""

    def sched(eta, k):
        return eta_min + (eta - eta_min) * 0.5 * (1.0 + math.cos(math.pi * k / max_iter))

    if warmup_iter == 0:
        logger.warning("warmup iters equals to zero, return CosineAnnealingLR")
    return ClosedFormLR(optimizer, sched, warmup_factor, warmup_iter, warmup_method)


def WarmupStepLR(optimizer, max_iter: int, warmup_factor: float, warmup_iter: int, step_size: int,
                 gamma: float = 0.1, warmup_method: str = "linear"):
    """``η·γ^⌊k/step_size⌋``."""

    def sched(eta, k):
        return eta * gamma ** (k // step_size)

    if warmup_iter == 0:
        logger.warning("warmup iters equals to zero, return StepLR")
    return ClosedFormLR(optimizer, sched, warmup_factor, warmup_iter, warmup_method)


def WarmupMultiStepLR(optimizer, max_iter: int, warmup_factor: float, warmup_iter: int, milestones: list,
                      gamma: float = 0.1, warmup_method: str = "linear"):
    """``η·γ^{#milestones ≤ k}``; milestones must be sorted and below ``max_iter``."""
    milestones = list(milestones)
    if milestones != sorted(milestones):
        raise ValueError(f"Milestones should be a list of increasing integers. Got {milestones}")
    if milestones and milestones[-1] > max_iter:
        raise ValueError(f"Milestones must be smaller than total training iterations {max_iter}. Got {milestones}")

    def sched(eta, k):
        return eta * gamma ** bisect.bisect_right(milestones, k)

    if warmup_iter == 0:
        logger.warning("warmup iters equals to zero, return MultiStepLR")
    return ClosedFormLR(optimizer, sched, warmup_factor, warmup_iter, warmup_method)


def WarmupExponentialLR(optimizer, max_iter: int, gamma: float, warmup_factor: float, warmup_iter: int,
                        warmup_method: str = "linear"):
    """``η·γ^k``."""

    def sched(eta, k):
        return eta * gamma ** k

    if warmup_iter == 0:
        logger.warning("warmup iters equals to zero, return ExponentialLR")
    return ClosedFormLR(optimizer, sched, warmup_factor, warmup_iter, warmup_method)


def WarmupPolynomialLR(optimizer, max_iter: int, warmup_factor: float, warmup_iter: int, end_learning_rate: float = 0.0001,
                       power: float = 1.0, cycle: bool = False, warmup_method: str = "linear"):
    """``(η−η_end)·(1−min(k,T)/T)^power + η_end`` with ``T = max_iter`` (``cycle`` stretches ``T``)."""

    def sched(eta, k):
        T = max_iter
        if cycle:
            T = T * max(1.0, math.ceil(k / T)) if k > 0 else T
            kk = k
        else:
            kk = min(k, T)
        return (eta - end_learning_rate) * (1.0 - kk / T) ** power + end_learning_rate

    if warmup_iter == 0:
        logger.warning("warmup iters equals to zero, return PolynomialLR")
    return ClosedFormLR(optimizer, sched, warmup_factor, warmup_iter, warmup_method)
