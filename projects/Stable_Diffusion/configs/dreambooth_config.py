"""DreamBooth: teach the model one subject from a handful of images (reference
projects/Stable_Diffusion/configs/dreambooth_config.py)."""
from transformers import CLIPTokenizer

from libai_b200.config import LazyCall
from projects.Stable_Diffusion.configs.config import dataloader, graph, model, optim, train  # noqa: F401
from projects.Stable_Diffusion.dataset import DreamBoothDataset

optim.lr = 5e-6
optim.weight_decay = 1e-2

dataloader.train.dataset = [
    LazyCall(DreamBoothDataset)(
        instance_data_root="/path/to/demo_dog/",
        instance_prompt="a photo of sks dog",
        tokenizer=CLIPTokenizer,
        tokenizer_pretrained_folder=["CompVis/stable-diffusion-v1-4", "tokenizer"],
    )
]

train.train_iter = 2000
train.log_period = 10
train.warmup_ratio = 0.0
train.output_dir = "output/stable_diffusion_dreambooth/"
