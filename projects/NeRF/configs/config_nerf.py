"""Train NeRF on a Blender or LLFF scene (reference projects/NeRF/configs/config_nerf.py).

    bash tools/train.sh tools/train_net.py projects/NeRF/configs/config_nerf.py 1 \
        train.dataset_type=Blender train.blender_dataset_path=/data/nerf_synthetic/lego

``train.dataset_type`` ∈ {Blender, LLFF, Analytic}; ``Analytic`` needs no files (procedural sphere).
``train.optim_type`` ∈ {adam, sgd, radam, ranger}; ``train.lr_scheduler_type`` ∈ {cosine, steplr}.
"""
import os

import torch

from libai_b200.config import LazyCall, get_config
from libai_b200.data.build import build_image_test_loader, build_image_train_loader
from libai_b200.optim import get_default_optimizer_params
from libai_b200.scheduler import WarmupCosineAnnealingLR, WarmupMultiStepLR
from projects.NeRF.configs.config_model import model
from projects.NeRF.datasets.nerf_dataset import get_nerf_dataset
from projects.NeRF.evaluation.nerf_evaluator import NerfEvaluator
from projects.NeRF.optimizers import RAdam, Ranger

graph = get_config("common/models/graph.py").graph
graph.enabled = False
train = get_config("common/train.py").train

train.dataset_type = os.environ.get("NERF_DATASET_TYPE", "Blender")     # Blender | LLFF | Analytic
train.blender_dataset_path = "/path/to/blender"
train.llff_dataset_path = "/path/to/llff"
train.optim_type = os.environ.get("NERF_OPTIM", "adam")
train.lr_scheduler_type = os.environ.get("NERF_SCHED", "cosine")
train.rays_per_batch = 1024          # rays per step and GPU
train.train_micro_batch_size = 1     # one item of the dataset already is a batch of rays
train.test_micro_batch_size = 1      # one image
train.train_epoch = {"Blender": 16, "LLFF": 30, "Analytic": 4}[train.dataset_type]
train.train_iter = 0
train.warmup_ratio = 0.0
train.evaluation.eval_period = 1000
train.log_period = 50

_img_wh = {"Blender": (400, 400), "LLFF": (504, 378), "Analytic": (48, 48)}[train.dataset_type]
_root = {"Blender": train.blender_dataset_path, "LLFF": train.llff_dataset_path, "Analytic": None}[train.dataset_type]

model.cfg.dataset_type = "Blender" if train.dataset_type == "Analytic" else train.dataset_type   # white background
model.cfg.loss_func = LazyCall(torch.nn.MSELoss)()
model.cfg.noise_std = 1.0 if train.dataset_type == "LLFF" else 0.0

train.evaluation = dict(
    enabled=True,
    evaluator=LazyCall(NerfEvaluator)(img_wh=_img_wh),
    eval_period=train.evaluation.eval_period,
    eval_iter=1e5,
    eval_metric="psnr",
    eval_mode="max",
)

_optimizers = {"adam": (torch.optim.Adam, 5e-4), "sgd": (torch.optim.SGD, 5e-2), "radam": (RAdam, 5e-4),
               "ranger": (Ranger, 5e-4)}
assert train.optim_type in _optimizers, "Nerf does not support this type of optimizer!"
_opt, _lr = _optimizers[train.optim_type]
optim = LazyCall(_opt)(
    params=LazyCall(get_default_optimizer_params)(clip_grad_max_norm=None, clip_grad_norm_type=None,
                                                  weight_decay_norm=None, weight_decay_bias=None),
    lr=_lr,
    weight_decay=0,
)
if train.optim_type == "sgd":
    optim.momentum = 0.9

if train.lr_scheduler_type == "steplr":
    train.scheduler = LazyCall(WarmupMultiStepLR)(
        warmup_factor=0.001, warmup_method="linear", gamma=0.5,
        milestones=[2 / 16, 4 / 16, 8 / 16] if train.dataset_type != "LLFF" else [10 / 30, 20 / 30],
    )
elif train.lr_scheduler_type == "cosine":
    train.scheduler = LazyCall(WarmupCosineAnnealingLR)(warmup_factor=0.001, warmup_method="linear", eta_min=1e-8)
else:
    raise NotImplementedError("Nerf does not support this type of scheduler!")

train.amp.enabled = True

dataset = LazyCall(get_nerf_dataset)(dataset_type=train.dataset_type)
_extra = dict(spheric_poses=False, val_num=1) if train.dataset_type == "LLFF" else {}

dataloader = dict(
    train=LazyCall(build_image_train_loader)(
        dataset=[LazyCall(dataset)(split="train", img_wh=_img_wh, root_dir=_root, batchsize=train.rays_per_batch,
                                   **_extra)],
        num_workers=0,        # the dataset keeps a sampling counter (centre-crop warm-up): keep it in-process
        train_batch_size=1,
        test_batch_size=train.test_micro_batch_size,
    ),
    test=[
        LazyCall(build_image_test_loader)(
            dataset=LazyCall(dataset)(split="val", img_wh=_img_wh, root_dir=_root, **_extra),
            num_workers=0,
            test_batch_size=train.test_micro_batch_size,
        )
    ],
)
from libai_b200.config import DictConfig  # noqa: E402

dataloader = DictConfig(dataloader)

train.dist.pipeline_num_layers = None
train.dist.data_parallel_size = 1
train.dist.tensor_parallel_size = 1
train.dist.pipeline_parallel_size = 1
