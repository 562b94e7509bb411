"""Swin-T on CIFAR-100, data parallel over 8 GPUs (reference configs/swin_cifar100.py)."""
from libai_b200.config import LazyCall
from libai_b200.data.vision import Mixup, SoftTargetCrossEntropy

from .common.data.cifar100 import dataloader
from .common.models.swin.swin_tiny_patch4_window7_224 import model
from .common.models.graph import graph
from .common.optim import optim
from .common.train import train

dataloader.train.mixup_func = LazyCall(Mixup)(
    mixup_alpha=0.8, cutmix_alpha=1.0, prob=1.0, switch_prob=0.5, mode="batch", num_classes=100
)
model.cfg.num_classes = 100
model.cfg.loss_func = SoftTargetCrossEntropy()

optim.lr = 5e-4
optim.eps = 1e-8
optim.weight_decay = 0.05
optim.params.clip_grad_max_norm = None
optim.params.clip_grad_norm_type = None

train.train_micro_batch_size = 32
train.num_accumulation_steps = 1
train.test_micro_batch_size = 32
train.train_epoch = 300
train.warmup_ratio = 20 / 300
train.evaluation.eval_period = 200
train.log_period = 20
train.scheduler.warmup_factor = 5e-7
train.scheduler.alpha = 0.0
train.scheduler.warmup_method = "linear"

train.dist.data_parallel_size = 8
train.dist.tensor_parallel_size = 1
train.dist.pipeline_parallel_size = 1
train.dist.pipeline_num_layers = sum(model.cfg.depths)
train.output_dir = "./output"

train.amp.enabled = False
train.activation_checkpoint.enabled = False
graph.enabled = False
