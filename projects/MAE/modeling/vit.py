"""ViT for MAE fine-tuning: global average pooling + ``fc_norm`` instead of the class token (reference
projects/MAE/modeling/vit.py)."""
from libai_b200.config import configurable
from libai_b200.layers import LayerNorm
from libai_b200.models import vision_transformer as core


class VisionTransformer(core.VisionTransformer):
    @configurable
    def __init__(self, *args, global_pool=False, **kwargs):
        super().__init__(*args, **kwargs)
        self.global_pool = global_pool
        if global_pool:
            self.fc_norm = LayerNorm(self.embed_dim if hasattr(self, "embed_dim") else self.norm.normalized_shape[0], layer_idx=-1)
            del self.norm

    @classmethod
    def from_config(cls, cfg):
        out = core.VisionTransformer.from_config.__func__(cls, cfg)
        out["global_pool"] = cfg.get("global_pool", False)
        return out

    def no_weight_decay(self):
        return {"pos_embed", "cls_token"}

    def forward_head(self, x):
        if self.global_pool:
            return self.head(self.fc_norm(x[:, 1:, :].mean(dim=1)))
        return super().forward_head(x)

    def forward_features(self, x):
        if not self.global_pool:
            return super().forward_features(x)
        import torch

        x = self.patch_embed(x)
        x = torch.cat((self.cls_token.to(x.dtype).expand(x.shape[0], -1, -1), x), dim=1)
        x = self.pos_drop(x + self.pos_embed.to(x.dtype))
        for blk in self.blocks:
            x = blk(x)
        return x
