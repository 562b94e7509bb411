"""CIFAR-100 loaders (recipe of reference configs/common/data/cifar100.py): 224-px random-resized crops, batch
Mixup/CutMix on the device."""
from torchvision import transforms
from torchvision.transforms import InterpolationMode

from libai_b200.config import LazyCall, OmegaConf
from libai_b200.data.build import build_image_test_loader, build_image_train_loader
from libai_b200.data.datasets import CIFAR100Dataset
from libai_b200.data.vision import Mixup, str_to_interp_mode

# channel statistics of the CIFAR-100 training split
CIFAR100_TRAIN_MEAN = (0.5070751592371323, 0.48654887331495095, 0.4409178433670343)
CIFAR100_TRAIN_STD = (0.2673342858792401, 0.2564384629170883, 0.27615047132568404)

_normalize = LazyCall(transforms.Normalize)(mean=CIFAR100_TRAIN_MEAN, std=CIFAR100_TRAIN_STD)

train_aug = LazyCall(transforms.Compose)(
    transforms=[
        LazyCall(transforms.RandomResizedCrop)(
            size=(224, 224), scale=(0.08, 1.0), ratio=(3.0 / 4.0, 4.0 / 3.0),
            interpolation=str_to_interp_mode("bicubic"),
        ),
        LazyCall(transforms.RandomHorizontalFlip)(),
        LazyCall(transforms.ToTensor)(),
        _normalize,
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
    dataset=[LazyCall(CIFAR100Dataset)(root="./", train=True, download=True, transform=train_aug)],
    num_workers=4,
    mixup_func=LazyCall(Mixup)(
        mixup_alpha=0.8, cutmix_alpha=1.0, prob=1.0, switch_prob=0.5, mode="batch", num_classes=100
    ),
)
dataloader.test = [
    LazyCall(build_image_test_loader)(
        dataset=LazyCall(CIFAR100Dataset)(root="./", train=False, download=True, transform=test_aug),
        num_workers=4,
    )
]
