"""Spec: reference libai/scheduler/build.py:19-23."""
from libai_b200.config import instantiate


def build_lr_scheduler(cfg, optimizer):
    """Instantiate the lazy scheduler record with ``optimizer`` injected."""
    cfg.optimizer = optimizer
    return instantiate(cfg)
