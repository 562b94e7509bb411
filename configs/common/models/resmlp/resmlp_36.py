from libai_b200.config import LazyCall
from libai_b200.models import ResMLP

from .resmlp_12 import cfg

cfg.depth = 36
cfg.init_scale = 1e-06

model = LazyCall(ResMLP)(cfg=cfg)
