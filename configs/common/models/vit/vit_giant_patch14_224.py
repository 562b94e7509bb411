from libai_b200.config import LazyCall
from libai_b200.models import VisionTransformer

from .vit_tiny_patch16_224 import cfg

cfg.patch_size = 14
cfg.embed_dim = 1408
cfg.mlp_ratio = 48 / 11
cfg.depth = 40
cfg.num_heads = 16

model = LazyCall(VisionTransformer)(cfg=cfg)
