"""ViT-Large/16 on synthetic 224×224 images (BASELINE.json config "ViT-Large configs/vit_imagenet.py DP=8 synthetic
224×224 images/sec"): the vit_imagenet recipe with ``vit_large_patch16_224`` and generated data."""
from libai_b200.config import LazyCall, OmegaConf
from libai_b200.data.build import build_image_test_loader, build_image_train_loader
from libai_b200.data.datasets import SyntheticImageDataset

from .common.models.vit.vit_large_patch16_224 import model
from .common.models.graph import graph
from .common.optim import optim
from .common.train import train

model.cfg.num_classes = 1000

dataloader = OmegaConf.create()
dataloader.train = LazyCall(build_image_train_loader)(
    dataset=[LazyCall(SyntheticImageDataset)(num_classes=1000, img_size=224, num_samples=1 << 20, seed=1234)],
    num_workers=4, mixup_func=None,
)
dataloader.test = [
    LazyCall(build_image_test_loader)(
        dataset=LazyCall(SyntheticImageDataset)(num_classes=1000, img_size=224, num_samples=256, seed=4321),
        num_workers=0,
    )
]

optim.lr = 1e-3
optim.weight_decay = 0.05
optim.params.clip_grad_max_norm = None
optim.params.clip_grad_norm_type = None
optim.params.overrides = {"pos_embed": {"weight_decay": 0.0}, "cls_token": {"weight_decay": 0.0}}

train.train_micro_batch_size = 128
train.test_micro_batch_size = 128
train.train_iter = 100
train.log_period = 10
train.amp.enabled = True
train.evaluation.enabled = False
train.dist.pipeline_num_layers = model.cfg.depth
train.output_dir = "./output/vit_large_synthetic"
