"""Stable Diffusion fine-tuning model: VAE encode → add noise → UNet predicts noise/velocity → MSE.

Spec: reference projects/Stable_Diffusion/modeling.py:29-135 — ``StableDiffusion(model_path, train_vae,
train_text_encoder, train_with_lora)`` with ``forward(pixel_values, input_ids) -> {"loss"}``; frozen VAE / text
encoder unless asked, LoRA mode freezing everything but the attention adapters, activation checkpointing of the UNet
and CLIP encoder blocks.

The UNet, VAE and scheduler are this repo's (``modules/``); tokenizer and text encoder are the ``transformers`` CLIP
classes the reference also uses.  ``model_path=None`` + ``tiny=True`` builds a small random-weight stack (unit
tests, smoke runs without a checkpoint).  With prior preservation (DreamBooth) the batch holds instance and class
images concatenated; the two halves are weighted by ``prior_loss_weight``.
"""
import torch
import torch.nn.functional as F
from torch import nn

from .modules.loader import load_submodel
from .modules.lora import add_lora_to_unet
from .modules.scheduler import DDPMScheduler
from .modules.unet import UNet2DConditionModel
from .modules.vae import AutoencoderKL

TINY_UNET = dict(block_out_channels=(32, 64), down_block_types=("CrossAttnDownBlock2D", "DownBlock2D"),
                 up_block_types=("UpBlock2D", "CrossAttnUpBlock2D"), layers_per_block=1, cross_attention_dim=32,
                 attention_head_dim=2, norm_num_groups=8)
TINY_VAE = dict(block_out_channels=(16, 32), layers_per_block=1, norm_num_groups=8)


def tiny_text_encoder(vocab_size=1000, hidden=32, max_len=16):
    from transformers import CLIPTextConfig, CLIPTextModel

    return CLIPTextModel(CLIPTextConfig(vocab_size=vocab_size, hidden_size=hidden, intermediate_size=64,
                                        num_hidden_layers=2, num_attention_heads=2, max_position_embeddings=max_len,
                                        projection_dim=hidden))


class StableDiffusion(nn.Module):
    def __init__(self, model_path=None, train_vae=False, train_text_encoder=False, train_with_lora=False,
                 lora_rank=4, with_prior_preservation=False, prior_loss_weight=1.0, tiny=False):
        super().__init__()
        self.model_path = model_path
        self.with_prior_preservation, self.prior_loss_weight = with_prior_preservation, prior_loss_weight
        if model_path is None:
            assert tiny, "give `model_path` (diffusers folder layout) or tiny=True for a random small model"
            self.tokenizer = None
            self.text_encoder = tiny_text_encoder()
            self.vae = AutoencoderKL(**TINY_VAE)
            self.unet = UNet2DConditionModel(**TINY_UNET)
            self.noise_scheduler = DDPMScheduler()
        else:
            from transformers import CLIPTextModel, CLIPTokenizer

            self.tokenizer = CLIPTokenizer.from_pretrained(model_path, subfolder="tokenizer")
            self.text_encoder = CLIPTextModel.from_pretrained(model_path, subfolder="text_encoder")
            self.vae = load_submodel(AutoencoderKL, model_path, "vae")
            self.unet = load_submodel(UNet2DConditionModel, model_path, "unet")
            self.noise_scheduler = DDPMScheduler.from_pretrained(model_path, subfolder="scheduler")
        self.scaling_factor = self.vae.config.get("scaling_factor", 0.18215)

        if not train_with_lora:
            if not train_vae:
                self.vae.requires_grad_(False)
            if not train_text_encoder:
                self.text_encoder.requires_grad_(False)
        else:
            self.vae.requires_grad_(False)
            self.text_encoder.requires_grad_(False)
            self.unet.requires_grad_(False)
            self.lora_layers = add_lora_to_unet(self.unet, rank=lora_rank)

    def train(self, mode=True):
        super().train(mode)
        for m in (self.vae, self.text_encoder):                       # frozen parts stay in eval mode
            if not any(p.requires_grad for p in m.parameters()):
                m.eval()
        return self

    def forward(self, pixel_values, input_ids):
        vae_grad = any(p.requires_grad for p in self.vae.parameters())
        with torch.set_grad_enabled(vae_grad and self.training):
            latents = self.vae.encode(pixel_values).latent_dist.sample() * self.scaling_factor
        noise = torch.randn_like(latents)
        bsz = latents.shape[0]
        timesteps = torch.randint(0, self.noise_scheduler.config.num_train_timesteps, (bsz,), device=latents.device)
        noisy_latents = self.noise_scheduler.add_noise(latents, noise, timesteps).to(self.unet.dtype)

        txt_grad = any(p.requires_grad for p in self.text_encoder.parameters())
        with torch.set_grad_enabled(txt_grad and self.training):
            encoder_hidden_states = self.text_encoder(input_ids)[0]

        noise_pred = self.unet(noisy_latents, timesteps, encoder_hidden_states).sample

        kind = self.noise_scheduler.config.prediction_type
        if kind == "epsilon":
            target = noise
        elif kind == "v_prediction":
            target = self.noise_scheduler.get_velocity(latents, noise, timesteps)
        else:
            raise ValueError(f"Unknown prediction type {kind}")

        if self.with_prior_preservation and bsz % 2 == 0:
            pred_inst, pred_prior = noise_pred.float().chunk(2, dim=0)
            tgt_inst, tgt_prior = target.float().chunk(2, dim=0)
            loss = F.mse_loss(pred_inst, tgt_inst) + self.prior_loss_weight * F.mse_loss(pred_prior, tgt_prior)
        else:
            loss = F.mse_loss(noise_pred.float(), target.float(), reduction="mean")
        return {"loss": loss}

    @staticmethod
    def set_activation_checkpoint(model):
        """Recompute the UNet blocks (and the CLIP encoder layers when it trains) in backward; the VAE is left
        alone (it usually runs without grad)."""
        model.unet.gradient_checkpointing = True
        if any(p.requires_grad for p in model.text_encoder.parameters()) and \
                hasattr(model.text_encoder, "gradient_checkpointing_enable"):
            model.text_encoder.gradient_checkpointing_enable()
