"""MoCo v3 ViT-S/16 pre-training (reference projects/MOCOV3/configs/moco_pretrain.py)."""
from torchvision import transforms

from libai_b200.config import LazyCall, get_config
from projects.MOCOV3.transform.pretrain_transform import TwoCropsTransform, augmentation1, augmentation2

from .models.moco_vit_small_patch16 import model

dataloader = get_config("common/data/imagenet.py").dataloader
train = get_config("common/train.py").train
graph = get_config("common/models/graph.py").graph
optim = get_config("common/optim.py").optim

dataloader.train.dataset[0].root = "/path/to/imagenet/"
dataloader.test[0].dataset.root = "/path/to/imagenet/"
dataloader.train.dataset[0].transform = LazyCall(TwoCropsTransform)(
    base_transform1=LazyCall(transforms.Compose)(transforms=augmentation1),
    base_transform2=LazyCall(transforms.Compose)(transforms=augmentation2),
)

model.m = 0.99   # momentum of the key encoder
model.T = 0.2    # softmax temperature

train.train_micro_batch_size = 32
train.test_micro_batch_size = 32
train.train_epoch = 300
train.warmup_ratio = 40 / 300
train.eval_period = 5
train.log_period = 1
train.evaluation.enabled = False

base_lr = 1.5e-4
optim.lr = base_lr * (train.train_micro_batch_size * 8 / 256)
optim.weight_decay = 0.1

train.scheduler.warmup_factor = 0.001
train.scheduler.alpha = 1.5e-4
train.scheduler.warmup_method = "linear"
graph.enabled = False
