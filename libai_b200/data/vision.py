"""Vision data utilities the reference takes from ``flowvision`` (configs/common/data/imagenet.py:1-30,
configs/vit_imagenet.py:7-8): RandAugment from a timm-style config string, random erasing, batch Mixup/CutMix with
label smoothing, and the soft-target cross entropy.  torchvision provides the image ops; Mixup runs on the device
on whole batches (``DefaultTrainer.get_batch``)."""
from __future__ import annotations

import math
import re
from typing import Optional, Tuple

import numpy as np
import torch
from torch import nn
from torchvision import transforms
from torchvision.transforms import InterpolationMode

IMAGENET_DEFAULT_MEAN = (0.485, 0.456, 0.406)
IMAGENET_DEFAULT_STD = (0.229, 0.224, 0.225)


def str_to_interp_mode(mode: str) -> InterpolationMode:
    return {"nearest": InterpolationMode.NEAREST, "bilinear": InterpolationMode.BILINEAR,
            "bicubic": InterpolationMode.BICUBIC, "lanczos": InterpolationMode.LANCZOS}[mode.lower()]


def rand_augment_transform(config_str: str, hparams: Optional[dict] = None):
    """``"rand-m9-mstd0.5-inc1"``: magnitude 9 (of 10 → 31 bins in torchvision), ``n`` ops per image (default 2).
    ``mstd`` / ``inc`` have no torchvision counterpart and are accepted for config compatibility."""
    hparams = hparams or {}
    parts = config_str.split("-")
    assert parts[0] == "rand", "only the 'rand-…' family is supported"
    magnitude, num_ops = 9, 2
    for part in parts[1:]:
        m = re.match(r"([a-z]+)([\d.]+)", part)
        if not m:
            continue
        key, value = m.group(1), float(m.group(2))
        if key == "m":
            magnitude = int(value)
        elif key == "n":
            num_ops = int(value)
    fill = hparams.get("img_mean")
    return transforms.RandAugment(
        num_ops=num_ops, magnitude=min(30, round(magnitude * 3)), num_magnitude_bins=31,
        interpolation=hparams.get("interpolation", InterpolationMode.BILINEAR),
        fill=list(fill) if fill is not None else None,
    )


class RandomErasing(transforms.RandomErasing):
    """``mode="pixel"`` erases with per-pixel noise, ``"const"`` with zeros; ``max_count`` rectangles."""

    def __init__(self, probability=0.5, min_area=0.02, max_area=1 / 3, min_aspect=0.3, mode="const", min_count=1,
                 max_count=None, num_splits=0, device="cpu"):
        super().__init__(p=probability, scale=(min_area, max_area), ratio=(min_aspect, 1 / min_aspect),
                         value="random" if mode == "pixel" else 0)
        self.count = max_count or min_count

    def forward(self, img):
        for _ in range(self.count):
            img = super().forward(img)
        return img


def _rand_bbox(h: int, w: int, lam: float, rng: np.random.RandomState) -> Tuple[int, int, int, int]:
    ratio = math.sqrt(1.0 - lam)
    ch, cw = int(h * ratio), int(w * ratio)
    cy, cx = rng.randint(h), rng.randint(w)
    return (max(cy - ch // 2, 0), min(cy + ch // 2, h), max(cx - cw // 2, 0), min(cx + cw // 2, w))


class Mixup:
    """Batch-mode Mixup / CutMix: ``(images [B,C,H,W], labels [B]) → (mixed images, soft targets [B, classes])``."""

    def __init__(self, mixup_alpha=1.0, cutmix_alpha=0.0, cutmix_minmax=None, prob=1.0, switch_prob=0.5, mode="batch",
                 correct_lam=True, label_smoothing=0.1, num_classes=1000, seed: Optional[int] = None):
        assert mode == "batch", "only batch mode is implemented"
        self.mixup_alpha, self.cutmix_alpha = mixup_alpha, cutmix_alpha
        self.prob, self.switch_prob = prob, switch_prob
        self.correct_lam, self.label_smoothing, self.num_classes = correct_lam, label_smoothing, num_classes
        self.rng = np.random.RandomState(seed)

    def _params(self):
        lam, use_cutmix = 1.0, False
        if self.rng.rand() < self.prob:
            if self.mixup_alpha > 0.0 and self.cutmix_alpha > 0.0:
                use_cutmix = self.rng.rand() < self.switch_prob
                alpha = self.cutmix_alpha if use_cutmix else self.mixup_alpha
            elif self.mixup_alpha > 0.0:
                alpha = self.mixup_alpha
            else:
                use_cutmix, alpha = True, self.cutmix_alpha
            lam = float(self.rng.beta(alpha, alpha))
        return lam, use_cutmix

    def _soft(self, target, lam):
        off = self.label_smoothing / self.num_classes
        on = 1.0 - self.label_smoothing + off
        y1 = torch.full((target.shape[0], self.num_classes), off, device=target.device).scatter_(1, target[:, None], on)
        return y1 * lam + y1.flip(0) * (1.0 - lam)

    def __call__(self, x, target):
        assert x.shape[0] % 2 == 0, "Batch size should be even when using this"
        lam, use_cutmix = self._params()
        if lam != 1.0:
            if use_cutmix:
                t, b, l, r = _rand_bbox(x.shape[-2], x.shape[-1], lam, self.rng)
                x = x.clone()
                x[:, :, t:b, l:r] = x.flip(0)[:, :, t:b, l:r]
                if self.correct_lam:
                    lam = 1.0 - (b - t) * (r - l) / float(x.shape[-2] * x.shape[-1])
            else:
                x = x * lam + x.flip(0) * (1.0 - lam)
        return x, self._soft(target, lam)


class SoftTargetCrossEntropy(nn.Module):
    def forward(self, x, target):
        return torch.sum(-target * torch.log_softmax(x.float(), dim=-1), dim=-1).mean()
