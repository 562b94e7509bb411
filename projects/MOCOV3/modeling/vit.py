"""ViT for MoCo v3 (reference projects/MOCOV3/modeling/vit.py): fixed 2-D sin-cos position embedding (class token
row zero), frozen random patch projection (``stop_grad_conv1``) for training stability, xavier-uniform q/k/v init
treating the fused matrix as three separate ones."""
import math

import torch
from torch import nn

from libai_b200.config import configurable
from libai_b200.layers import Linear
from libai_b200.models import vision_transformer as core


class VisionTransformer(core.VisionTransformer):
    @configurable
    def __init__(self, *args, stop_grad_conv1=False, linear_prob=None, **kwargs):
        super().__init__(*args, **kwargs)
        self.stop_grad_conv1, self.linear_prob = stop_grad_conv1, linear_prob
        self.initialization()

    @classmethod
    def from_config(cls, cfg):
        out = core.VisionTransformer.from_config.__func__(cls, cfg)
        out["stop_grad_conv1"] = cfg.get("stop_grad_conv1", False)
        out["linear_prob"] = cfg.get("linear_prob", None)
        return out

    def initialization(self):
        with torch.no_grad():
            dim = self.pos_embed.shape[-1]
            grid = int(math.sqrt(self.pos_embed.shape[1] - 1))
            self.pos_embed.copy_(self.build_2d_sincos_position_embedding(grid, dim))
            self.pos_embed.requires_grad = False
            for name, m in self.named_modules():
                if isinstance(m, Linear) and m.weight.device.type != "meta":
                    if "query_key_value" in name:
                        val = math.sqrt(6.0 / float(m.weight.shape[0] // 3 + m.weight.shape[1]))
                        m.weight.uniform_(-val, val)
                    else:
                        nn.init.xavier_uniform_(m.weight)
                    if m.bias is not None:
                        m.bias.zero_()
            nn.init.normal_(self.cls_token, std=1e-6)
            proj = self.patch_embed.proj
            w = proj.weight
            if w.device.type != "meta":
                fan = w.shape[1] if w.dim() == 2 else w[0].numel()
                val = math.sqrt(6.0 / float(fan + w.shape[0]))
                w.uniform_(-val, val)
                if getattr(proj, "bias", None) is not None:
                    proj.bias.zero_()
            if self.stop_grad_conv1:
                for p in self.patch_embed.proj.parameters():
                    p.requires_grad = False

    @staticmethod
    def build_2d_sincos_position_embedding(grid, embed_dim, temperature=10000.0):
        assert embed_dim % 4 == 0, "Embed dimension must be divisible by 4 for 2D sin-cos position embedding"
        gw, gh = torch.meshgrid(torch.arange(grid, dtype=torch.float32), torch.arange(grid, dtype=torch.float32), indexing="ij")
        pos_dim = embed_dim // 4
        omega = 1.0 / (temperature ** (torch.arange(pos_dim, dtype=torch.float32) / pos_dim))
        out_w, out_h = torch.einsum("m,d->md", gw.flatten(), omega), torch.einsum("m,d->md", gh.flatten(), omega)
        emb = torch.cat([out_w.sin(), out_w.cos(), out_h.sin(), out_h.cos()], dim=1)[None]
        return torch.cat([torch.zeros(1, 1, embed_dim), emb], dim=1)

    def forward_head(self, x):
        if self.linear_prob:
            x = x.detach()  # linear probing: only the head trains
        return super().forward_head(x)
