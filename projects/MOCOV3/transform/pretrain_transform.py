"""MoCo v3 / BYOL augmentations (reference projects/MOCOV3/transform/pretrain_transform.py): two differently
augmented crops stacked on the channel axis."""
import random

import torch
from PIL import ImageFilter, ImageOps
from torchvision import transforms

from libai_b200.config import LazyCall
from libai_b200.data.vision import IMAGENET_DEFAULT_MEAN, IMAGENET_DEFAULT_STD


class GaussianBlur:
    def __init__(self, sigma=(0.1, 2.0)):
        self.sigma = sigma

    def __call__(self, x):
        return x.filter(ImageFilter.GaussianBlur(radius=random.uniform(self.sigma[0], self.sigma[1])))


class Solarize:
    def __call__(self, x):
        return ImageOps.solarize(x)


def _aug(blur_p, solarize_p):
    return [
        LazyCall(transforms.RandomResizedCrop)(size=224, scale=(0.2, 1.0)),
        LazyCall(transforms.RandomApply)(transforms=[LazyCall(transforms.ColorJitter)(brightness=0.4, contrast=0.4, saturation=0.2, hue=0.1)], p=0.8),
        LazyCall(transforms.RandomGrayscale)(p=0.2),
        LazyCall(transforms.RandomApply)(transforms=[LazyCall(GaussianBlur)(sigma=[0.1, 2.0])], p=blur_p),
        LazyCall(transforms.RandomApply)(transforms=[LazyCall(Solarize)()], p=solarize_p),
        LazyCall(transforms.RandomHorizontalFlip)(),
        LazyCall(transforms.ToTensor)(),
        LazyCall(transforms.Normalize)(mean=IMAGENET_DEFAULT_MEAN, std=IMAGENET_DEFAULT_STD),
    ]


augmentation1 = _aug(1.0, 0.0)
augmentation2 = _aug(0.1, 0.2)


class TwoCropsTransform:
    def __init__(self, base_transform1, base_transform2):
        self.base_transform1, self.base_transform2 = base_transform1, base_transform2

    def __call__(self, x):
        return torch.cat((self.base_transform1(x), self.base_transform2(x)), dim=0)
