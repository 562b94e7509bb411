"""Linear probing on frozen MoCo v3 features (reference projects/MOCOV3/configs/moco_linear_prob.py)."""
from libai_b200.config import LazyCall, get_config
from libai_b200.optim import SGD
from projects.MOCOV3.transform.linear_prob_transform import test_aug, train_aug

from .models.vit_small_patch16 import model

dataloader = get_config("common/data/imagenet.py").dataloader
train = get_config("common/train.py").train
graph = get_config("common/models/graph.py").graph
optim = get_config("common/optim.py").optim

dataloader.train.dataset[0].root = "/path/to/imagenet"
dataloader.test[0].dataset.root = "/path/to/imagenet"
dataloader.train.dataset[0].transform = train_aug
dataloader.test[0].dataset.transform = test_aug

model.linear_prob = dict(weight_style="oneflow", path="/path/to/moco/checkpoint")

train.train_micro_batch_size = 128
train.test_micro_batch_size = 32
train.train_epoch = 90
train.warmup_ratio = 0
train.eval_period = 1000
train.log_period = 100

optim._target_ = SGD
optim.lr = 0.1 * (train.train_micro_batch_size * 8 / 256)
optim.weight_decay = 0.0
optim.momentum = 0.9
for _k in ("betas", "eps", "do_bias_correction"):
    optim.pop(_k, None)

train.scheduler.warmup_factor = 0.001
train.scheduler.alpha = 0.0
train.scheduler.warmup_method = "linear"
graph.enabled = False
