"""SwinIR image restoration / super-resolution network (Liang et al. 2021), used to upsample the 64×64 DALL-E 2
samples 4× / 16×.

Spec: reference projects/DALLE2/swinir/models.py:17-1035.  Structure: shallow conv → N residual Swin-transformer
blocks (RSTB: a stack of (shifted-)window attention layers + conv, with a long skip) → conv → upsampler
(``pixelshuffle`` | ``pixelshuffledirect`` | ``nearest+conv``).  The window-attention core is shared with this
framework's Swin classifier (``libai_b200/models/swin_transformer.py``); parameter names follow the public SwinIR
checkpoints so the released ``*.pth`` files load with ``strict=True``.  Unlike the classifier the resolution is not
fixed: attention masks are built (and cached) per input size.
"""
import math

import torch
import torch.nn.functional as F
from torch import nn

from libai_b200.layers import Linear
from libai_b200.models.swin_transformer import WindowAttention, shifted_window_mask, window_partition, window_reverse


class Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = Linear(dim, hidden)
        self.fc2 = Linear(hidden, dim)

    def forward(self, x):
        return self.fc2(self.fc1(x, "gelu"))


class SwinLayer(nn.Module):
    def __init__(self, dim, num_heads, window_size, shift_size, mlp_ratio, qkv_bias=True):
        super().__init__()
        self.window_size, self.shift_size = window_size, shift_size
        self.norm1 = nn.LayerNorm(dim)
        self.attn = WindowAttention(dim, (window_size, window_size), num_heads, qkv_bias=qkv_bias)
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))
        self._masks = {}

    def _mask(self, H, W, device):
        if self.shift_size == 0:
            return None
        key = (H, W, str(device))
        if key not in self._masks:
            self._masks[key] = shifted_window_mask(H, W, self.window_size, self.shift_size, device)
        return self._masks[key]

    def forward(self, x, x_size):
        H, W = x_size
        B, L, C = x.shape
        ws, ss = self.window_size, self.shift_size
        h = self.norm1(x).view(B, H, W, C)
        if ss > 0:
            h = torch.roll(h, shifts=(-ss, -ss), dims=(1, 2))
        win = window_partition(h, ws).view(-1, ws * ws, C)
        win = self.attn(win, self._mask(H, W, x.device)).view(-1, ws, ws, C)
        h = window_reverse(win, ws, H, W)
        if ss > 0:
            h = torch.roll(h, shifts=(ss, ss), dims=(1, 2))
        x = x + h.reshape(B, L, C)
        return x + self.mlp(self.norm2(x))


class _Blocks(nn.Module):
    def __init__(self, dim, depth, num_heads, window_size, mlp_ratio, qkv_bias):
        super().__init__()
        self.blocks = nn.ModuleList([SwinLayer(dim, num_heads, window_size, 0 if i % 2 == 0 else window_size // 2,
                                               mlp_ratio, qkv_bias) for i in range(depth)])

    def forward(self, x, x_size):
        for b in self.blocks:
            x = b(x, x_size)
        return x


class RSTB(nn.Module):
    """Residual Swin Transformer Block: attention layers → (un-embed) conv (re-embed) → + input."""

    def __init__(self, dim, depth, num_heads, window_size, mlp_ratio, qkv_bias, resi_connection):
        super().__init__()
        self.residual_group = _Blocks(dim, depth, num_heads, window_size, mlp_ratio, qkv_bias)
        if resi_connection == "1conv":
            self.conv = nn.Conv2d(dim, dim, 3, 1, 1)
        else:       # "3conv": bottleneck saves parameters in the large models
            self.conv = nn.Sequential(nn.Conv2d(dim, dim // 4, 3, 1, 1), nn.LeakyReLU(0.2, inplace=True),
                                      nn.Conv2d(dim // 4, dim // 4, 1, 1, 0), nn.LeakyReLU(0.2, inplace=True),
                                      nn.Conv2d(dim // 4, dim, 3, 1, 1))

    def forward(self, x, x_size):
        B, L, C = x.shape
        y = self.residual_group(x, x_size)
        y = self.conv(y.transpose(1, 2).reshape(B, C, *x_size))
        return x + y.flatten(2).transpose(1, 2)


class _Norm(nn.Module):
    """``patch_embed`` of SwinIR: flatten + LayerNorm (the name keeps checkpoint keys ``patch_embed.norm.*``)."""

    def __init__(self, dim):
        super().__init__()
        self.norm = nn.LayerNorm(dim)

    def forward(self, x):
        return self.norm(x.flatten(2).transpose(1, 2))


class Upsample(nn.Sequential):
    def __init__(self, scale, num_feat):
        layers = []
        if (scale & (scale - 1)) == 0:
            for _ in range(int(math.log2(scale))):
                layers += [nn.Conv2d(num_feat, 4 * num_feat, 3, 1, 1), nn.PixelShuffle(2)]
        elif scale == 3:
            layers += [nn.Conv2d(num_feat, 9 * num_feat, 3, 1, 1), nn.PixelShuffle(3)]
        else:
            raise ValueError(f"scale {scale} is not supported. Supported scales: 2^n and 3.")
        super().__init__(*layers)


class UpsampleOneStep(nn.Sequential):
    def __init__(self, scale, num_feat, num_out_ch):
        super().__init__(nn.Conv2d(num_feat, (scale ** 2) * num_out_ch, 3, 1, 1), nn.PixelShuffle(scale))


class SwinIR(nn.Module):
    def __init__(self, img_size=64, patch_size=1, in_chans=3, embed_dim=96, depths=(6, 6, 6, 6), num_heads=(6, 6, 6, 6),
                 window_size=7, mlp_ratio=4.0, qkv_bias=True, upscale=2, img_range=1.0, upsampler="",
                 resi_connection="1conv", **unused):
        super().__init__()
        num_feat = 64
        self.img_range, self.upscale, self.upsampler, self.window_size = img_range, upscale, upsampler, window_size
        mean = torch.tensor((0.4488, 0.4371, 0.4040)).view(1, 3, 1, 1) if in_chans == 3 else torch.zeros(1, 1, 1, 1)
        self.register_buffer("mean", mean, persistent=False)
        self.conv_first = nn.Conv2d(in_chans, embed_dim, 3, 1, 1)
        self.patch_embed = _Norm(embed_dim)
        self.layers = nn.ModuleList([RSTB(embed_dim, d, h, window_size, mlp_ratio, qkv_bias, resi_connection)
                                     for d, h in zip(depths, num_heads)])
        self.norm = nn.LayerNorm(embed_dim)
        if resi_connection == "1conv":
            self.conv_after_body = nn.Conv2d(embed_dim, embed_dim, 3, 1, 1)
        else:
            self.conv_after_body = nn.Sequential(nn.Conv2d(embed_dim, embed_dim // 4, 3, 1, 1), nn.LeakyReLU(0.2, inplace=True),
                                                 nn.Conv2d(embed_dim // 4, embed_dim // 4, 1, 1, 0), nn.LeakyReLU(0.2, inplace=True),
                                                 nn.Conv2d(embed_dim // 4, embed_dim, 3, 1, 1))
        if upsampler == "pixelshuffle":
            self.conv_before_upsample = nn.Sequential(nn.Conv2d(embed_dim, num_feat, 3, 1, 1), nn.LeakyReLU(inplace=True))
            self.upsample = Upsample(upscale, num_feat)
            self.conv_last = nn.Conv2d(num_feat, in_chans, 3, 1, 1)
        elif upsampler == "pixelshuffledirect":
            self.upsample = UpsampleOneStep(upscale, embed_dim, in_chans)
        elif upsampler == "nearest+conv":
            assert upscale == 4, "only support x4 now."
            self.conv_before_upsample = nn.Sequential(nn.Conv2d(embed_dim, num_feat, 3, 1, 1), nn.LeakyReLU(inplace=True))
            self.conv_up1 = nn.Conv2d(num_feat, num_feat, 3, 1, 1)
            self.conv_up2 = nn.Conv2d(num_feat, num_feat, 3, 1, 1)
            self.conv_hr = nn.Conv2d(num_feat, num_feat, 3, 1, 1)
            self.conv_last = nn.Conv2d(num_feat, in_chans, 3, 1, 1)
            self.lrelu = nn.LeakyReLU(negative_slope=0.2, inplace=True)
        else:       # denoising / JPEG artefact removal
            self.conv_last = nn.Conv2d(embed_dim, in_chans, 3, 1, 1)

    def check_image_size(self, x):
        _, _, h, w = x.shape
        ws = self.window_size
        return F.pad(x, (0, (ws - w % ws) % ws, 0, (ws - h % ws) % ws), mode="reflect")

    def forward_features(self, x):
        x_size = (x.shape[2], x.shape[3])
        t = self.patch_embed(x)
        for layer in self.layers:
            t = layer(t, x_size)
        t = self.norm(t)
        return t.transpose(1, 2).reshape(x.shape[0], -1, *x_size)

    def load_state_dict(self, state_dict, strict=True, **kw):
        # released checkpoints carry the (input-size specific) attention masks; they are rebuilt on the fly here
        state_dict = {k: v for k, v in state_dict.items() if not k.endswith("attn_mask")}
        return super().load_state_dict(state_dict, strict=strict, **kw)

    def forward(self, x):
        H, W = x.shape[2:]
        x = self.check_image_size(x)
        mean = self.mean.to(x.dtype)
        x = (x - mean) * self.img_range
        if self.upsampler == "pixelshuffle":
            x = self.conv_first(x)
            x = self.conv_after_body(self.forward_features(x)) + x
            x = self.conv_last(self.upsample(self.conv_before_upsample(x)))
        elif self.upsampler == "pixelshuffledirect":
            x = self.conv_first(x)
            x = self.upsample(self.conv_after_body(self.forward_features(x)) + x)
        elif self.upsampler == "nearest+conv":
            x = self.conv_first(x)
            x = self.conv_after_body(self.forward_features(x)) + x
            x = self.conv_before_upsample(x)
            x = self.lrelu(self.conv_up1(F.interpolate(x, scale_factor=2, mode="nearest")))
            x = self.lrelu(self.conv_up2(F.interpolate(x, scale_factor=2, mode="nearest")))
            x = self.conv_last(self.lrelu(self.conv_hr(x)))
        else:
            first = self.conv_first(x)
            x = x + self.conv_last(self.conv_after_body(self.forward_features(first)) + first)
        x = x / self.img_range + mean
        return x[:, :, : H * self.upscale, : W * self.upscale]
