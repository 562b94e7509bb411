"""MAE fine-tuning of ViT-B/16 with layer-wise LR decay, Mixup/CutMix (reference projects/MAE/configs/mae_finetune.py)."""
from libai_b200.config import LazyCall, get_config
from libai_b200.data.vision import Mixup, SoftTargetCrossEntropy
from projects.MAE.configs.models.vit_base_patch16 import model
from projects.MAE.utils.lr_decay import param_groups_lrd
from projects.MAE.utils.scheduler import warmup_layerscale_cosine_lr_scheduler

train = get_config("common/train.py").train
optim = get_config("common/optim.py").optim
graph = get_config("common/models/graph.py").graph
dataloader = get_config("common/data/imagenet.py").dataloader

graph.enabled = False
dataloader.train.dataset[0].root = "/path/to/imagenet"
dataloader.test[0].dataset.root = "/path/to/imagenet"
dataloader.train.mixup_func = LazyCall(Mixup)(mixup_alpha=0.8, cutmix_alpha=1.0, prob=1.0, switch_prob=0.5,
                                              mode="batch", label_smoothing=0.1, num_classes=1000)

finetune = dict(enable=True, weight_style="oneflow", path="/path/to/pretrained_mae_weight")  # or "pytorch"
model.loss_func = LazyCall(SoftTargetCrossEntropy)()

n_gpus = 8
train.train_micro_batch_size = 32
train.num_accumulation_steps = 4
train.test_micro_batch_size = 32
effective_batch_size = train.train_micro_batch_size * train.num_accumulation_steps * n_gpus
train.train_epoch = 100
train.warmup_ratio = 5 / 100
train.log_period = 20
train.evaluation.eval_after_n_epoch = 1
train.checkpointer.save_model_after_n_epoch = 1

base_lr = 5e-4
actual_lr = base_lr * effective_batch_size / 256
optim.params._target_ = param_groups_lrd
optim.params.weight_decay = 0.05
optim.params.layer_decay = 0.65
optim.lr = actual_lr
for _k in ("clip_grad_max_norm", "clip_grad_norm_type", "weight_decay_norm", "weight_decay_bias"):
    optim.params.pop(_k, None)
optim.pop("weight_decay", None)

train.scheduler = LazyCall(warmup_layerscale_cosine_lr_scheduler)(warmup_factor=0.0, min_lr=1e-6)
train.amp.enabled = True
train.dist.data_parallel_size = n_gpus
train.dist.tensor_parallel_size = 1
train.dist.pipeline_parallel_size = 1
