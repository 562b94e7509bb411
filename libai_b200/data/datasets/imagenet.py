"""ImageNet-1k folder dataset returning ``Instance(images, labels)`` (spec: reference
libai/data/datasets/imagenet.py:25-50; flowvision is replaced by torchvision)."""
import os
from typing import Callable, Optional

import torch
from torchvision import datasets

from libai_b200.data.structures import DistTensorData, Instance


class ImageNetDataset(datasets.ImageFolder):
    """``root/{train,val}/<class>/*.JPEG``; ``train`` selects the split."""

    def __init__(self, root: str, train: bool = True, transform: Optional[Callable] = None, **kwargs):
        super().__init__(root=os.path.join(root, "train" if train else "val"), transform=transform, **kwargs)

    def __getitem__(self, index: int):
        sample, target = super().__getitem__(index)
        return Instance(
            images=DistTensorData(sample, placement_idx=0),
            labels=DistTensorData(torch.tensor(target, dtype=torch.long), placement_idx=-1),
        )
