"""CIFAR-10 / CIFAR-100 as ``Instance(images, labels)`` (spec: reference libai/data/datasets/cifar.py:24-96)."""
from typing import Callable, Optional

import torch
from torchvision import datasets

from libai_b200.data.structures import DistTensorData, Instance


class _InstanceMixin:
    def __getitem__(self, index: int):
        img, target = super().__getitem__(index)
        return Instance(
            images=DistTensorData(img, placement_idx=0),
            labels=DistTensorData(torch.tensor(target, dtype=torch.long), placement_idx=-1),
        )


class CIFAR10Dataset(_InstanceMixin, datasets.CIFAR10):
    def __init__(self, root: str, train: bool = True, transform: Optional[Callable] = None, download: bool = False, **kwargs):
        super().__init__(root=root, train=train, transform=transform, download=download, **kwargs)


class CIFAR100Dataset(_InstanceMixin, datasets.CIFAR100):
    def __init__(self, root: str, train: bool = True, transform: Optional[Callable] = None, download: bool = False, **kwargs):
        super().__init__(root=root, train=train, transform=transform, download=download, **kwargs)
