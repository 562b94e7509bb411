"""MAE ViT-B/16 pre-training, 800 epochs on 8 GPUs (reference projects/MAE/configs/mae_pretraining.py)."""
from torchvision import transforms
from torchvision.transforms import InterpolationMode

from libai_b200.config import LazyCall, get_config
from libai_b200.data.vision import IMAGENET_DEFAULT_MEAN, IMAGENET_DEFAULT_STD
from projects.MAE.configs.models.mae_vit_base_patch16 import model
from projects.MAE.data.pretraining_imagenet import PretrainingImageNetDataset
from projects.MAE.utils.lr_decay import param_groups_weight_decay
from projects.MAE.utils.scheduler import warmup_cosine_lr_scheduler

train = get_config("common/train.py").train
optim = get_config("common/optim.py").optim
graph = get_config("common/models/graph.py").graph
dataloader = get_config("common/data/imagenet.py").dataloader

graph.enabled = True
dataloader.train.dataset[0].root = "/path/to/imagenet"
dataloader.train.dataset[0]._target_ = PretrainingImageNetDataset
del dataloader.test  # no evaluation during pre-training

dataloader.train.dataset[0].transform = LazyCall(transforms.Compose)(
    transforms=[
        LazyCall(transforms.RandomResizedCrop)(size=(224, 224), scale=(0.2, 1.0), interpolation=InterpolationMode.BICUBIC),
        LazyCall(transforms.RandomHorizontalFlip)(),
        LazyCall(transforms.ToTensor)(),
        LazyCall(transforms.Normalize)(mean=IMAGENET_DEFAULT_MEAN, std=IMAGENET_DEFAULT_STD),
    ]
)

n_gpus = 8
train.train_micro_batch_size = 64
train.num_accumulation_steps = 8
effective_batch_size = train.train_micro_batch_size * train.num_accumulation_steps * n_gpus
train.train_epoch = 800
train.warmup_ratio = 40 / 800
train.log_period = 20
train.checkpointer.save_model_after_n_epoch = 20

base_lr = 1.5e-4
actual_lr = base_lr * effective_batch_size / 256  # linear scaling rule

optim.params._target_ = param_groups_weight_decay
optim.params.weight_decay = 0.05
optim.lr = actual_lr
optim.betas = (0.9, 0.95)
for _k in ("clip_grad_max_norm", "clip_grad_norm_type", "weight_decay_norm", "weight_decay_bias"):
    optim.params.pop(_k, None)
optim.pop("weight_decay", None)

train.scheduler = LazyCall(warmup_cosine_lr_scheduler)(warmup_factor=0.0, min_lr=0.0)
train.amp.enabled = True
train.evaluation.enabled = False
train.dist.data_parallel_size = n_gpus
train.dist.tensor_parallel_size = 1
train.dist.pipeline_parallel_size = 1
