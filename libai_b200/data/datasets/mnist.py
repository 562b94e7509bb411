"""MNIST as ``Instance(images, labels)`` (spec: reference libai/data/datasets/mnist.py:24-58)."""
from typing import Callable, Optional

import torch
from torchvision import datasets

from libai_b200.data.structures import DistTensorData, Instance


class MNISTDataset(datasets.MNIST):
    def __init__(self, root: str, train: bool = True, transform: Optional[Callable] = None, download: bool = False, **kwargs):
        super().__init__(root=root, train=train, transform=transform, download=download, **kwargs)

    def __getitem__(self, index: int):
        img, target = super().__getitem__(index)
        return Instance(
            images=DistTensorData(img, placement_idx=0),
            labels=DistTensorData(torch.tensor(int(target), dtype=torch.long), placement_idx=-1),
        )
