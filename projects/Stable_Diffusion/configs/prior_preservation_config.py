"""DreamBooth with prior preservation: every batch pairs instance images with generated class images (see
``generate_prior_image.py``) so the class does not collapse onto the subject (reference
projects/Stable_Diffusion/configs/prior_preservation_config.py)."""
from projects.Stable_Diffusion.configs.dreambooth_config import dataloader, graph, model, optim, train  # noqa: F401
from projects.Stable_Diffusion.dataset import prior_preservation_collate

dataloader.train.dataset[0].class_data_root = "/path/to/prior_dog/"
dataloader.train.dataset[0].class_prompt = "a photo of dog"
dataloader.train.collate_fn = prior_preservation_collate

model.with_prior_preservation = True
model.prior_loss_weight = 1.0

train.output_dir = "output/stable_diffusion_prior_preservation/"
