"""Warm-up schedulers vs closed forms (model: reference tests/test_scheduler.py)."""
import math

import torch

from libai_b200.scheduler import (
    WarmupCosineLR,
    WarmupExponentialLR,
    WarmupMultiStepLR,
    WarmupPolynomialLR,
    WarmupStepLR,
)


def _run(sched_fn, n):
    p = torch.nn.Parameter(torch.zeros(2))
    opt = torch.optim.SGD([p], lr=5.0)
    sched = sched_fn(opt)
    p.sum().backward()
    lrs = []
    for _ in range(n):
        lrs.append(opt.param_groups[0]["lr"])
        opt.step()
        sched.step()
    return lrs


def test_warmup_multistep():
    lrs = _run(lambda o: WarmupMultiStepLR(o, max_iter=30, milestones=[10, 15, 20], gamma=0.1, warmup_factor=0.001,
                                           warmup_iter=5, warmup_method="linear"), 30)
    assert all(abs(a - b) < 1e-9 for a, b in zip(lrs[:5], [0.005, 1.004, 2.003, 3.002, 4.001]))
    assert all(abs(x - 5.0) < 1e-9 for x in lrs[5:10])
    assert all(abs(x - 0.5) < 1e-9 for x in lrs[10:15])
    assert all(abs(x - 0.05) < 1e-9 for x in lrs[15:20])
    assert all(abs(x - 0.005) < 1e-9 for x in lrs[20:])


def test_warmup_cosine():
    lrs = _run(lambda o: WarmupCosineLR(o, max_iter=30, warmup_factor=0.001, warmup_iter=5, warmup_method="linear"), 30)
    # warm-up ramps linearly from warmup_factor * base_lr to the cosine value at the end of the warm-up
    assert abs(lrs[0] - 0.005) < 1e-9 and all(lrs[i] < lrs[i + 1] for i in range(5))
    for idx in range(5, 30):
        expected = 2.5 * (1.0 + math.cos(math.pi * idx / 30))
        assert abs(lrs[idx] - expected) < 1e-6, (idx, lrs[idx], expected)


def test_warmup_exponential():
    lrs = _run(lambda o: WarmupExponentialLR(o, max_iter=10, gamma=0.1, warmup_factor=0.001, warmup_iter=5,
                                             warmup_method="linear"), 10)
    assert abs(lrs[0] - 0.005) < 1e-9
    for idx in range(5, 10):
        assert abs(lrs[idx] - 5 * 0.1 ** idx) < 1e-9


def test_warmup_step():
    lrs = _run(lambda o: WarmupStepLR(o, max_iter=30, step_size=10, gamma=0.1, warmup_factor=0.001, warmup_iter=5,
                                      warmup_method="linear"), 30)
    assert all(abs(x - 5.0) < 1e-9 for x in lrs[5:10]) and all(abs(x - 0.5) < 1e-9 for x in lrs[10:20])
    assert all(abs(x - 0.05) < 1e-9 for x in lrs[20:])


def test_warmup_polynomial():
    lrs = _run(lambda o: WarmupPolynomialLR(o, max_iter=30, warmup_factor=0.001, warmup_iter=0, end_learning_rate=1e-4,
                                            power=1.0, cycle=False), 30)
    for idx in range(30):
        assert abs(lrs[idx] - ((5.0 - 1e-4) * (1 - idx / 30) + 1e-4)) < 1e-6
