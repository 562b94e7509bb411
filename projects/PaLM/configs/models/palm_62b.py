from .palm_small import model

model.cfg.dim = 8192
model.cfg.depth = 64
model.cfg.dim_head = 256
model.cfg.num_heads = 32
