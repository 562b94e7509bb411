"""Swin-T on ImageNet-1k (reference configs/swin_imagenet.py)."""
from libai_b200.config import LazyCall
from libai_b200.data.vision import Mixup, SoftTargetCrossEntropy

from .common.data.imagenet import dataloader
from .common.models.swin.swin_tiny_patch4_window7_224 import model
from .common.models.graph import graph
from .common.optim import optim
from .common.train import train

dataloader.train.mixup_func = LazyCall(Mixup)(
    mixup_alpha=0.8, cutmix_alpha=1.0, prob=1.0, switch_prob=0.5, mode="batch", num_classes=1000
)
model.cfg.num_classes = 1000
model.cfg.loss_func = SoftTargetCrossEntropy()

dataloader.train.dataset[0].root = "/path/to/imagenet"
dataloader.test[0].dataset.root = "/path/to/imagenet"

optim.lr = 1e-3
optim.eps = 1e-8
optim.weight_decay = 0.05
optim.params.clip_grad_max_norm = None
optim.params.clip_grad_norm_type = None

train.train_micro_batch_size = 128
train.test_micro_batch_size = 128
train.train_epoch = 300
train.warmup_ratio = 20 / 300
train.eval_period = 1562
train.log_period = 100
train.scheduler.warmup_factor = 0.001
train.scheduler.alpha = 0.01
train.scheduler.warmup_method = "linear"
train.amp.enabled = True
