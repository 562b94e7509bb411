from .palm_small import model

model.cfg.dim = 6144
model.cfg.depth = 48
model.cfg.dim_head = 256
model.cfg.num_heads = 24
