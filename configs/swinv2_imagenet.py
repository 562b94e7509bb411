"""SwinV2-T (window 8, 256 px) on ImageNet-1k (reference configs/swinv2_imagenet.py)."""
from libai_b200.config import LazyCall
from libai_b200.data.vision import Mixup, SoftTargetCrossEntropy
from libai_b200.optim import set_weight_decay

from .common.data.imagenet import dataloader
from .common.models.swinv2.swinv2_tiny_patch4_window8_256 import model
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

# 256-px pipeline instead of the default 224-px one
dataloader.train.dataset[0].transform.transforms[0].size = 256
dataloader.train.dataset[0].transform.transforms[2].hparams.translate_const = int(256 * 0.45)
dataloader.test[0].dataset.transform.transforms[0].size = 256
dataloader.test[0].dataset.transform.transforms[1].size = 256

optim.lr = 1e-3
optim.eps = 1e-8
optim.weight_decay = 0.05
# no decay for 1-D tensors / biases / the position-bias machinery
optim.params = LazyCall(set_weight_decay)(
    model=model,
    skip_list=("absolute_pos_embed",),
    skip_keywords=("cpb_mlp", "logit_scale", "relative_position_bias_table"),
)

train.train_micro_batch_size = 128
train.test_micro_batch_size = 128
train.train_epoch = 300
train.warmup_ratio = 20 / 300
train.eval_period = 1562
train.log_period = 100
graph.enabled = False
train.rdma_enabled = True
train.scheduler.warmup_factor = 0.001
train.scheduler.alpha = 0.01
train.scheduler.warmup_method = "linear"
train.amp.enabled = True
