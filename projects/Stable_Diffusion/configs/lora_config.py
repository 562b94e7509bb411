"""LoRA fine-tuning: only rank-4 adapters on the UNet attention projections train (reference
projects/Stable_Diffusion/configs/lora_config.py)."""
from projects.Stable_Diffusion.configs.dreambooth_config import dataloader, graph, model, optim, train  # noqa: F401

optim.lr = 5e-4
model.train_with_lora = True
model.lora_rank = 4

train.train_iter = 2000
train.zero_optimization.enabled = False       # a few MB of trainable parameters: nothing to shard
train.output_dir = "output/stable_diffusion_lora/"
