"""ImageNet-1k loaders: RandomResizedCrop + flip + RandAugment(m9) + RandomErasing(0.25) for training,
Resize(256)/CenterCrop(224) for evaluation (recipe of reference configs/common/data/imagenet.py; torchvision ops)."""
from torchvision import transforms
from torchvision.transforms import InterpolationMode

from libai_b200.config import LazyCall, OmegaConf
from libai_b200.data.build import build_image_test_loader, build_image_train_loader
from libai_b200.data.datasets import ImageNetDataset
from libai_b200.data.vision import (
    IMAGENET_DEFAULT_MEAN,
    IMAGENET_DEFAULT_STD,
    RandomErasing,
    rand_augment_transform,
    str_to_interp_mode,
)

_normalize = LazyCall(transforms.Normalize)(mean=IMAGENET_DEFAULT_MEAN, std=IMAGENET_DEFAULT_STD)

train_aug = LazyCall(transforms.Compose)(
    transforms=[
        LazyCall(transforms.RandomResizedCrop)(
            size=224, scale=(0.08, 1.0), ratio=(3.0 / 4.0, 4.0 / 3.0), interpolation=InterpolationMode.BICUBIC
        ),
        LazyCall(transforms.RandomHorizontalFlip)(p=0.5),
        LazyCall(rand_augment_transform)(
            config_str="rand-m9-mstd0.5-inc1",
            hparams=dict(
                translate_const=int(224 * 0.45),
                img_mean=tuple(min(255, round(255 * x)) for x in IMAGENET_DEFAULT_MEAN),
                interpolation=str_to_interp_mode("bicubic"),
            ),
        ),
        LazyCall(transforms.ToTensor)(),
        _normalize,
        LazyCall(RandomErasing)(probability=0.25, mode="pixel", max_count=1, num_splits=0, device="cpu"),
    ]
)

test_aug = LazyCall(transforms.Compose)(
    transforms=[
        LazyCall(transforms.Resize)(size=256, interpolation=InterpolationMode.BICUBIC),
        LazyCall(transforms.CenterCrop)(size=224),
        LazyCall(transforms.ToTensor)(),
        _normalize,
    ]
)

dataloader = OmegaConf.create()
dataloader.train = LazyCall(build_image_train_loader)(
    dataset=[LazyCall(ImageNetDataset)(root="./dataset", train=True, transform=train_aug)],
    num_workers=4,
    mixup_func=None,
)
dataloader.test = [
    LazyCall(build_image_test_loader)(
        dataset=LazyCall(ImageNetDataset)(root="./dataset", train=False, transform=test_aug),
        num_workers=4,
    )
]
