"""Text → image sampling with classifier-free guidance (the inference side the reference delegates to
``onediff.OneFlowStableDiffusionPipeline``, projects/Stable_Diffusion/train_net.py:26,60-72 and
generate_prior_image.py).  ``save_pretrained`` / ``from_pretrained`` use the diffusers folder layout."""
import os

import torch

from .modules.loader import load_submodel, save_submodel
from .modules.lora import load_attn_procs
from .modules.scheduler import DDPMScheduler
from .modules.unet import UNet2DConditionModel
from .modules.vae import AutoencoderKL


class StableDiffusionPipeline:
    def __init__(self, tokenizer, text_encoder, vae, unet, scheduler):
        self.tokenizer, self.text_encoder, self.vae, self.unet, self.scheduler = tokenizer, text_encoder, vae, unet, scheduler

    @classmethod
    def from_pretrained(cls, model_path, tokenizer=None, text_encoder=None, vae=None, unet=None, dtype=None):
        from transformers import CLIPTextModel, CLIPTokenizer

        tokenizer = tokenizer or CLIPTokenizer.from_pretrained(model_path, subfolder="tokenizer")
        text_encoder = text_encoder or CLIPTextModel.from_pretrained(model_path, subfolder="text_encoder")
        vae = vae or load_submodel(AutoencoderKL, model_path, "vae")
        unet = unet or load_submodel(UNet2DConditionModel, model_path, "unet")
        pipe = cls(tokenizer, text_encoder, vae, unet, DDPMScheduler.from_pretrained(model_path))
        return pipe.to(dtype=dtype) if dtype is not None else pipe

    def save_pretrained(self, save_dir):
        import json

        os.makedirs(save_dir, exist_ok=True)
        save_submodel(self.unet, save_dir, "unet", "UNet2DConditionModel")
        save_submodel(self.vae, save_dir, "vae", "AutoencoderKL")
        os.makedirs(os.path.join(save_dir, "scheduler"), exist_ok=True)
        with open(os.path.join(save_dir, "scheduler", "scheduler_config.json"), "w") as f:
            json.dump({"_class_name": "DDPMScheduler", **self.scheduler.config}, f, indent=2)
        self.text_encoder.save_pretrained(os.path.join(save_dir, "text_encoder"))
        if self.tokenizer is not None:
            self.tokenizer.save_pretrained(os.path.join(save_dir, "tokenizer"))

    def load_lora(self, path, scale=1.0):
        return load_attn_procs(self.unet, path, scale=scale)

    def to(self, device=None, dtype=None):
        for m in (self.text_encoder, self.vae, self.unet):
            m.to(device=device, dtype=dtype)
        return self

    def _encode(self, prompts, device):
        tok = self.tokenizer(prompts, padding="max_length", max_length=self.tokenizer.model_max_length,
                             truncation=True, return_tensors="pt")
        return self.text_encoder(tok.input_ids.to(device))[0]

    @torch.no_grad()
    def __call__(self, prompt=None, *, input_ids=None, negative_input_ids=None, height=512, width=512,
                 num_inference_steps=50, guidance_scale=7.5, eta=0.0, generator=None, output_type="pil"):
        device = next(self.unet.parameters()).device
        for m in (self.text_encoder, self.vae, self.unet):
            m.eval()
        if input_ids is None:
            prompts = [prompt] if isinstance(prompt, str) else list(prompt)
            cond = self._encode(prompts, device)
            uncond = self._encode([""] * len(prompts), device)
        else:
            cond = self.text_encoder(input_ids.to(device))[0]
            neg = negative_input_ids if negative_input_ids is not None else torch.zeros_like(input_ids)
            uncond = self.text_encoder(neg.to(device))[0]
        b = cond.shape[0]
        scale = 2 ** (len(self.vae.config["block_out_channels"]) - 1)
        latents = torch.randn(b, self.unet.config["in_channels"], height // scale, width // scale,
                              generator=generator, device=device if generator is None else generator.device)
        latents = latents.to(device=device, dtype=self.unet.dtype)
        self.scheduler.set_timesteps(num_inference_steps, device=device)
        ctx = torch.cat([uncond, cond], dim=0) if guidance_scale > 1.0 else cond
        for t in self.scheduler.timesteps:
            inp = torch.cat([latents, latents], dim=0) if guidance_scale > 1.0 else latents
            out = self.unet(inp, t, ctx).sample
            if guidance_scale > 1.0:
                u, c = out.chunk(2, dim=0)
                out = u + guidance_scale * (c - u)
            latents = self.scheduler.step_ddim(out, t, latents, eta=eta, generator=None)
        images = self.vae.decode(latents / self.vae.config.get("scaling_factor", 0.18215)).sample
        images = (images.float() / 2 + 0.5).clamp(0, 1)
        if output_type == "pt":
            return images
        from PIL import Image

        arr = (images.permute(0, 2, 3, 1).cpu().numpy() * 255).round().astype("uint8")
        return [Image.fromarray(a) for a in arr]
