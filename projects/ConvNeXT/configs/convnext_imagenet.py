"""ConvNeXt-T on ImageNet-1k (reference projects/ConvNeXT/configs/convnext_imagenet.py)."""
from configs.common.data.imagenet import dataloader
from configs.common.models.graph import graph
from configs.common.optim import optim
from configs.common.train import train
from libai_b200.config import LazyCall
from libai_b200.data.vision import Mixup
from projects.ConvNeXT.configs.convnext import model

dataloader.train.dataset[0].root = "/data/dataset/ImageNet/extract"
dataloader.test[0].dataset.root = "/data/dataset/ImageNet/extract"
model.cfg.num_labels = 1000
dataloader.train.mixup_func = LazyCall(Mixup)(
    mixup_alpha=0.8, cutmix_alpha=1.0, prob=1.0, switch_prob=0.5, mode="batch", num_classes=model.cfg.num_labels
)

optim.lr = 1e-3  # 5e-4 * 1024 (batch) / 512
optim.eps = 1e-8
optim.weight_decay = 0.05
optim.params.clip_grad_max_norm = None
optim.params.clip_grad_norm_type = None

train.train_micro_batch_size = 128
train.test_micro_batch_size = 128
train.train_epoch = 300
train.warmup_ratio = 5 / 300
train.evaluation.eval_period = 1000
train.log_period = 1
train.scheduler.warmup_factor = 0.001
train.scheduler.alpha = 0.01
train.scheduler.warmup_method = "linear"
train.amp.enabled = True
train.dist.data_parallel_size = 1
train.dist.tensor_parallel_size = 1
train.dist.pipeline_parallel_size = 1
