"""Uniform view of a CLIP model for the prior and decoder (reference projects/DALLE2/dalle2/_clip.py:22-116):
``embed_text`` → (pooled embedding, per-token encodings, token mask); ``embed_image`` → (pooled embedding, None).
Backed by this repo's ``projects/CLIP`` implementation."""
import torch
import torch.nn.functional as F
from torch import nn


class OpenAIClipAdapter(nn.Module):
    def __init__(self, name="ViT-L/14", clip=None, download_root=None):
        super().__init__()
        if clip is None:
            from projects.CLIP.clip import load

            clip, _ = load(name, device="cpu", download_root=download_root)
        self.clip = clip
        self.eos_id = 49407
        self.register_buffer("mean", torch.tensor([0.48145466, 0.4578275, 0.40821073]).view(1, 3, 1, 1), persistent=False)
        self.register_buffer("std", torch.tensor([0.26862954, 0.26130258, 0.27577711]).view(1, 3, 1, 1), persistent=False)

    @property
    def dim_latent(self):
        return self.clip.text_projection.shape[-1]

    @property
    def image_size(self):
        return self.clip.visual.input_resolution

    @property
    def image_channels(self):
        return 3

    @property
    def max_text_len(self):
        return self.clip.context_length

    @torch.no_grad()
    def embed_text(self, text):
        c = self.clip
        text = text[..., : self.max_text_len]
        # everything up to and including the first EOT token is real text
        is_eos = text == self.eos_id
        eos_pos = torch.where(is_eos.any(dim=-1), is_eos.float().argmax(dim=-1), text.argmax(dim=-1))
        mask = torch.arange(text.shape[1], device=text.device)[None] <= eos_pos[:, None]
        x = c.token_embedding(text).type(c.dtype) + c.positional_embedding.type(c.dtype)
        x = c.transformer(x.permute(1, 0, 2)).permute(1, 0, 2)
        enc = c.ln_final(x).type(c.dtype)
        pooled = enc[torch.arange(enc.shape[0], device=text.device), eos_pos] @ c.text_projection
        enc = enc.masked_fill(~mask[..., None], 0.0)
        return F.normalize(pooled.float(), dim=-1), enc.float(), mask

    @torch.no_grad()
    def embed_image(self, image):
        image = F.interpolate(image, size=(self.image_size, self.image_size), mode="bicubic", align_corners=False)
        image = (image - self.mean) / self.std
        return F.normalize(self.clip.encode_image(image).float(), dim=-1), None
