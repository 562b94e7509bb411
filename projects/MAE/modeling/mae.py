"""Masked Autoencoder with a ViT backbone.

Spec: reference projects/MAE/modeling/mae.py — ``MaskedAutoencoderViT`` (:34-400): patch embedding + fixed 2-D
sin-cos positions, per-sample random masking by arg-sorting noise (:242-268), encoder on the kept 25 %, light decoder
with mask tokens restored by ``ids_restore`` (:306-350), MSE on the masked patches with optional per-patch
normalisation (:352-373).  Blocks are the library ``TransformerLayer`` (native LayerNorm / GEMM / MLP kernels; the
encoder's short, variable-length sequences use the flash kernel when the shape allows).
"""
import torch
from torch import nn

from libai_b200.config import configurable
from libai_b200.layers import LayerNorm, Linear, PatchEmbedding, TransformerLayer
from libai_b200.layers._param import create_parameter, trunc_normal_, zeros_

from .pos_embed import get_2d_sincos_pos_embed


def _xavier(t, generator=None):
    fan_out, fan_in = t.shape[0], t[0].numel()
    bound = (6.0 / (fan_in + fan_out)) ** 0.5
    return t.uniform_(-bound, bound, generator=generator)


class MaskedAutoencoderViT(nn.Module):
    @configurable
    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=1024, depth=24, num_heads=16,
                 decoder_embed_dim=512, decoder_depth=8, decoder_num_heads=16, mlp_ratio=4.0, norm_layer=LayerNorm,
                 norm_pix_loss=False, mask_ratio=0.75):
        super().__init__()
        self.mask_ratio, self.norm_pix_loss, self.in_chans = mask_ratio, norm_pix_loss, in_chans
        self.patch_embed = PatchEmbedding(img_size, patch_size, in_chans, embed_dim)
        num_patches = self.patch_embed.num_patches
        self.cls_token = create_parameter((1, 1, embed_dim), lambda t, generator=None: t.normal_(std=0.02, generator=generator))
        self.register_buffer("pos_embed", torch.from_numpy(
            get_2d_sincos_pos_embed(embed_dim, int(num_patches ** 0.5), cls_token=True)).float()[None], persistent=True)
        self.blocks = nn.ModuleList([
            TransformerLayer(embed_dim, int(embed_dim * mlp_ratio), num_heads, init_method=_xavier, layer_idx=i)
            for i in range(depth)])
        self.norm = norm_layer(embed_dim, layer_idx=depth - 1)
        # decoder
        self.decoder_embed = Linear(embed_dim, decoder_embed_dim, bias=True, init_method=_xavier, layer_idx=depth)
        self.mask_token = create_parameter((1, 1, decoder_embed_dim), lambda t, generator=None: t.normal_(std=0.02, generator=generator))
        self.register_buffer("decoder_pos_embed", torch.from_numpy(
            get_2d_sincos_pos_embed(decoder_embed_dim, int(num_patches ** 0.5), cls_token=True)).float()[None], persistent=True)
        self.decoder_blocks = nn.ModuleList([
            TransformerLayer(decoder_embed_dim, int(decoder_embed_dim * mlp_ratio), decoder_num_heads, init_method=_xavier,
                             layer_idx=depth + i) for i in range(decoder_depth)])
        self.decoder_norm = norm_layer(decoder_embed_dim, layer_idx=-1)
        self.decoder_pred = Linear(decoder_embed_dim, patch_size ** 2 * in_chans, bias=True, init_method=_xavier, layer_idx=-1)
        self.patch_size = patch_size

    @classmethod
    def from_config(cls, cfg):
        keys = ("img_size patch_size in_chans embed_dim depth num_heads decoder_embed_dim decoder_depth decoder_num_heads "
                "mlp_ratio norm_pix_loss mask_ratio").split()
        return {k: cfg[k] for k in keys if k in cfg}

    def patchify(self, imgs):
        """``[N, C, H, W]`` → ``[N, L, p²·C]``."""
        p, c = self.patch_size, imgs.shape[1]
        h = w = imgs.shape[2] // p
        x = imgs.reshape(imgs.shape[0], c, h, p, w, p)
        return torch.einsum("nchpwq->nhwpqc", x).reshape(imgs.shape[0], h * w, p * p * c)

    def unpatchify(self, x):
        p, c = self.patch_size, self.in_chans
        h = w = int(x.shape[1] ** 0.5)
        x = x.reshape(x.shape[0], h, w, p, p, c)
        return torch.einsum("nhwpqc->nchpwq", x).reshape(x.shape[0], c, h * p, w * p)

    @staticmethod
    def random_masking(x, mask_ratio):
        n, length, dim = x.shape
        keep = int(length * (1 - mask_ratio))
        noise = torch.rand(n, length, device=x.device)
        ids_shuffle = torch.argsort(noise, dim=1)
        ids_restore = torch.argsort(ids_shuffle, dim=1)
        ids_keep = ids_shuffle[:, :keep]
        x_masked = torch.gather(x, 1, ids_keep[..., None].expand(-1, -1, dim))
        mask = torch.ones(n, length, device=x.device)
        mask[:, :keep] = 0
        return x_masked, torch.gather(mask, 1, ids_restore), ids_restore

    def forward_encoder(self, x, mask_ratio):
        x = self.patch_embed(x)
        x = x + self.pos_embed[:, 1:].to(x.dtype)
        x, mask, ids_restore = self.random_masking(x, mask_ratio)
        cls = (self.cls_token + self.pos_embed[:, :1]).to(x.dtype).expand(x.shape[0], -1, -1)
        x = torch.cat((cls, x), dim=1)
        for blk in self.blocks:
            x = blk(x)
        return self.norm(x), mask, ids_restore

    def forward_decoder(self, x, ids_restore):
        x = self.decoder_embed(x)
        n_mask = ids_restore.shape[1] + 1 - x.shape[1]
        tokens = torch.cat([x[:, 1:], self.mask_token.to(x.dtype).expand(x.shape[0], n_mask, -1)], dim=1)
        tokens = torch.gather(tokens, 1, ids_restore[..., None].expand(-1, -1, x.shape[2]))
        x = torch.cat([x[:, :1], tokens], dim=1) + self.decoder_pos_embed.to(x.dtype)
        for blk in self.decoder_blocks:
            x = blk(x)
        return self.decoder_pred(self.decoder_norm(x))[:, 1:]

    def forward_loss(self, imgs, pred, mask):
        target = self.patchify(imgs).float()
        if self.norm_pix_loss:
            target = (target - target.mean(-1, keepdim=True)) / (target.var(-1, keepdim=True) + 1e-6) ** 0.5
        loss = ((pred.float() - target) ** 2).mean(-1)
        return (loss * mask).sum() / mask.sum()

    def forward(self, images):
        latent, mask, ids_restore = self.forward_encoder(images, self.mask_ratio)
        pred = self.forward_decoder(latent, ids_restore)
        loss = self.forward_loss(images, pred, mask)
        if self.training:
            return {"losses": loss}
        return {"losses": loss, "pred": pred, "mask": mask}
