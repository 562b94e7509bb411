"""SwinV2-T (window 8, 256 px) on CIFAR-100 (reference configs/swinv2_cifar100.py)."""
from libai_b200.config import LazyCall
from libai_b200.data.vision import Mixup, SoftTargetCrossEntropy
from libai_b200.data.vision import str_to_interp_mode
from libai_b200.optim import set_weight_decay
from torchvision import transforms
from torchvision.transforms import InterpolationMode

from .common.data.cifar100 import CIFAR100_TRAIN_MEAN, CIFAR100_TRAIN_STD
from .common.data.cifar100 import dataloader
from .common.models.swinv2.swinv2_tiny_patch4_window8_256 import model
from .common.models.graph import graph
from .common.optim import optim
from .common.train import train

dataloader.train.mixup_func = LazyCall(Mixup)(
    mixup_alpha=0.8, cutmix_alpha=1.0, prob=1.0, switch_prob=0.5, mode="batch", num_classes=100
)
model.cfg.num_classes = 100
model.cfg.loss_func = SoftTargetCrossEntropy()

_normalize = LazyCall(transforms.Normalize)(mean=CIFAR100_TRAIN_MEAN, std=CIFAR100_TRAIN_STD)
dataloader.train.dataset[0].transform = LazyCall(transforms.Compose)(
    transforms=[
        LazyCall(transforms.RandomResizedCrop)(
            size=(256, 256), scale=(0.08, 1.0), ratio=(3.0 / 4.0, 4.0 / 3.0),
            interpolation=str_to_interp_mode("bicubic"),
        ),
        LazyCall(transforms.RandomHorizontalFlip)(),
        LazyCall(transforms.ToTensor)(),
        _normalize,
    ]
)
dataloader.test[0].dataset.transform = LazyCall(transforms.Compose)(
    transforms=[
        LazyCall(transforms.Resize)(size=256, interpolation=InterpolationMode.BICUBIC),
        LazyCall(transforms.CenterCrop)(size=256),
        LazyCall(transforms.ToTensor)(),
        _normalize,
    ]
)

optim.lr = 5e-4
optim.eps = 1e-8
optim.weight_decay = 0.05
# no decay for 1-D tensors / biases / the position-bias machinery
optim.params = LazyCall(set_weight_decay)(
    model=model,
    skip_list=("absolute_pos_embed",),
    skip_keywords=("cpb_mlp", "logit_scale", "relative_position_bias_table"),
)

train.train_micro_batch_size = 32
train.num_accumulation_steps = 8
train.test_micro_batch_size = 32
train.train_epoch = 300
train.warmup_ratio = 20 / 300
train.evaluation.eval_period = 1562
train.log_period = 10
train.scheduler.warmup_factor = 5e-7
train.scheduler.alpha = 0.0
train.scheduler.warmup_method = "linear"

train.dist.data_parallel_size = 1
train.dist.tensor_parallel_size = 1
train.dist.pipeline_parallel_size = 1
train.dist.pipeline_num_layers = sum(model.cfg.depths)
train.output_dir = "./output"
train.rdma_enabled = False

train.amp.enabled = False
train.activation_checkpoint.enabled = False
graph.enabled = False
