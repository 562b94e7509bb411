"""Linear-probing augmentations (reference projects/MOCOV3/transform/linear_prob_transform.py)."""
from torchvision import transforms

from libai_b200.config import LazyCall
from libai_b200.data.vision import IMAGENET_DEFAULT_MEAN, IMAGENET_DEFAULT_STD

_norm = LazyCall(transforms.Normalize)(mean=IMAGENET_DEFAULT_MEAN, std=IMAGENET_DEFAULT_STD)
train_aug = LazyCall(transforms.Compose)(transforms=[
    LazyCall(transforms.RandomResizedCrop)(size=224), LazyCall(transforms.RandomHorizontalFlip)(),
    LazyCall(transforms.ToTensor)(), _norm])
test_aug = LazyCall(transforms.Compose)(transforms=[
    LazyCall(transforms.Resize)(size=256), LazyCall(transforms.CenterCrop)(size=224), LazyCall(transforms.ToTensor)(), _norm])
