"""ResMLP (https://arxiv.org/abs/2105.03404).

Spec: reference libai/models/resmlp.py:31-295 — ``Affine`` (α·x+β), residual blocks with a
cross-patch linear (on the transposed token axis) and a per-patch MLP, both scaled by learnable
``gamma`` (layer scale, init ``init_scale``), Affine "norm", mean-pool head; DP/TP/PP capable.
"""
from __future__ import annotations

import torch
from torch import nn

from libai_b200.config import configurable
from libai_b200.layers import MLP, DropPath, Linear, PatchEmbedding
from libai_b200.layers._param import create_parameter, ones_, trunc_normal_, zeros_

from .utils.pipeline_model import PipelineStageMixin


def _tn(t, generator=None):
    return trunc_normal_(t, std=0.02, generator=generator)


class Affine(nn.Module):
    def __init__(self, dim, *, layer_idx=0):
        super().__init__()
        self.alpha = create_parameter((dim,), ones_, layer_idx=layer_idx)
        self.beta = create_parameter((dim,), zeros_, layer_idx=layer_idx)
        self.layer_idx = layer_idx

    def forward(self, x):
        return self.alpha.to(x.dtype) * x + self.beta.to(x.dtype)


class layers_scale_mlp_blocks(nn.Module):
    def __init__(self, dim, drop=0.0, drop_path=0.0, init_values=1e-4, num_patches=196, *, layer_idx=0):
        super().__init__()
        self.norm1 = Affine(dim, layer_idx=layer_idx)
        self.attn = Linear(num_patches, num_patches, init_method=_tn, layer_idx=layer_idx)
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()
        self.norm2 = Affine(dim, layer_idx=layer_idx)
        self.mlp = MLP(hidden_size=dim, ffn_hidden_size=int(4.0 * dim), init_method=_tn, layer_idx=layer_idx)

        def scale_init(t, generator=None):
            return t.fill_(init_values)

        self.gamma_1 = create_parameter((dim,), scale_init, layer_idx=layer_idx)
        self.gamma_2 = create_parameter((dim,), scale_init, layer_idx=layer_idx)
        self.layer_idx = layer_idx

    def forward(self, x):
        mixed = self.attn(self.norm1(x).transpose(1, 2).contiguous()).transpose(1, 2)
        x = x + self.drop_path(self.gamma_1.to(x.dtype) * mixed)
        x = x + self.drop_path(self.gamma_2.to(x.dtype) * self.mlp(self.norm2(x)))
        return x


class ResMLP(nn.Module, PipelineStageMixin):
    @configurable
    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, depth=12, drop_rate=0.0,
                 drop_path_rate=0.0, init_scale=1e-4, num_classes=1000, loss_func=None):
        super().__init__()
        self.num_classes = num_classes
        self.num_features = self.embed_dim = embed_dim
        self.patch_embed = PatchEmbedding(img_size=img_size, patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim)
        num_patches = self.patch_embed.num_patches
        self.blocks = nn.ModuleList(
            [
                layers_scale_mlp_blocks(dim=embed_dim, drop=drop_rate, drop_path=drop_path_rate, init_values=init_scale,
                                        num_patches=num_patches, layer_idx=i)
                for i in range(depth)
            ]
        )
        self.norm = Affine(embed_dim, layer_idx=-1)
        self.head = Linear(embed_dim, num_classes, init_method=_tn, layer_idx=-1) if num_classes > 0 else nn.Identity()
        self.loss_func = nn.CrossEntropyLoss() if loss_func is None else loss_func

    @classmethod
    def from_config(cls, cfg):
        keys = "img_size patch_size in_chans embed_dim depth drop_rate drop_path_rate init_scale num_classes loss_func".split()
        return {k: cfg[k] for k in keys}

    def stage_pre(self, images, **_):
        return self.patch_embed(images)

    def stage_layers(self):
        return self.blocks

    def stage_post(self, hidden, labels=None, **_):
        x = self.forward_head(hidden)
        if labels is not None and self.training:
            return {"losses": self.loss_func(x.float(), labels)}
        return {"prediction_scores": x}

    def forward_features(self, x):
        x = self.patch_embed(x)
        for blk in self.blocks:
            x = blk(x)
        return x

    def forward_head(self, x):
        return self.head(self.norm(x).mean(dim=1))

    def forward(self, images, labels=None):
        return self.forward_stage({"images": images, "labels": labels})

    @staticmethod
    def set_pipeline_stage_id(model):
        return model

    @staticmethod
    def set_activation_checkpoint(model):
        model.activation_checkpoint = True
        return model
