"""ResMLP-12 on ImageNet-1k with LAMB (reference configs/resmlp_imagenet.py)."""
from libai_b200.config import LazyCall
from libai_b200.data.vision import Mixup, SoftTargetCrossEntropy
from libai_b200.optim import LAMB

from .common.data.imagenet import dataloader
from .common.models.resmlp.resmlp_12 import model
from .common.models.graph import graph
from .common.optim import optim
from .common.train import train

dataloader.train.mixup_func = LazyCall(Mixup)(
    mixup_alpha=0.8, cutmix_alpha=1.0, prob=1.0, switch_prob=0.5, mode="batch", num_classes=1000
)
model.cfg.num_classes = 1000
model.cfg.loss_func = SoftTargetCrossEntropy()

train.output_dir = "./output_resmlp"
dataloader.train.dataset[0].root = "/path/to/imagenet"
dataloader.test[0].dataset.root = "/path/to/imagenet"

# ResMLP evaluates on a 224/0.9 resize
dataloader.test[0].dataset.transform.transforms[0].size = int(224 / 0.9)

optim._target_ = LAMB
optim.lr = 5e-3  # global batch 256 * 8 = 2048
optim.eps = 1e-8
optim.weight_decay = 0.2
optim.params.clip_grad_max_norm = None
optim.params.clip_grad_norm_type = None
optim.params.overrides = {
    "alpha": {"weight_decay": 0.0},
    "beta": {"weight_decay": 0.0},
    "gamma_1": {"weight_decay": 0.0},
    "gamma_2": {"weight_decay": 0.0},
}

train.train_micro_batch_size = 256
train.test_micro_batch_size = 64
train.train_epoch = 400
train.warmup_ratio = 5 / 400
train.evaluation.eval_period = 1000
train.log_period = 1
train.scheduler.warmup_factor = 0.001
train.scheduler.alpha = 0.01
train.scheduler.warmup_method = "linear"
train.amp.enabled = True

train.dist.pipeline_num_layers = model.cfg.depth
train.dist.data_parallel_size = 1
train.dist.tensor_parallel_size = 1
train.dist.pipeline_parallel_size = 1
