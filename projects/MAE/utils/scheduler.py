"""Half-cycle cosine schedule with linear warm-up that honours per-group ``lr_scale`` (layer-wise decay) — reference
projects/MAE/utils/scheduler.py."""
import math

from torch.optim.lr_scheduler import _LRScheduler


class LayerScaleWarmupCosineDecayLR(_LRScheduler):
    def __init__(self, optimizer, steps, warmup_steps, warmup_factor=0.0, min_lr=0.0, last_step=-1, verbose=False):
        self.total_steps, self.warmup_steps, self.warmup_factor, self.min_lr = steps, warmup_steps, warmup_factor, min_lr
        super().__init__(optimizer, last_step)

    def _lr(self, base_lr, step):
        if step < self.warmup_steps:
            return base_lr * (self.warmup_factor + (1 - self.warmup_factor) * step / max(1, self.warmup_steps))
        progress = (step - self.warmup_steps) / max(1, self.total_steps - self.warmup_steps)
        return self.min_lr + (base_lr - self.min_lr) * 0.5 * (1.0 + math.cos(math.pi * min(1.0, progress)))

    def get_lr(self):
        return [self._lr(b, self.last_epoch) * g.get("lr_scale", 1.0) for b, g in zip(self.base_lrs, self.optimizer.param_groups)]


def warmup_layerscale_cosine_lr_scheduler(optimizer, max_iter, warmup_iter, warmup_factor, min_lr=0.0, **kwargs):
    return LayerScaleWarmupCosineDecayLR(optimizer, steps=max_iter, warmup_steps=warmup_iter, warmup_factor=warmup_factor, min_lr=min_lr)


def warmup_cosine_lr_scheduler(optimizer, max_iter, warmup_iter, warmup_factor, min_lr=0.0, **kwargs):
    return LayerScaleWarmupCosineDecayLR(optimizer, steps=max_iter, warmup_steps=warmup_iter, warmup_factor=warmup_factor, min_lr=min_lr)
