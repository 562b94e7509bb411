"""KL-regularised convolutional autoencoder (``AutoencoderKL``): images ↔ 8×-downsampled 4-channel latents.

Parameter names follow the diffusers checkpoint layout.  The reference takes this model from ``diffusers``
(projects/Stable_Diffusion/modeling.py:39)."""
import torch
import torch.nn.functional as F
from torch import nn

from libai_b200.ops import functional as OF

from .unet import Downsample2D, ResnetBlock2D, Upsample2D


class VaeAttention(nn.Module):
    """Single-head spatial self-attention of the VAE mid block."""

    def __init__(self, channels, groups=32):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, channels, eps=1e-6)
        self.to_q = nn.Linear(channels, channels)
        self.to_k = nn.Linear(channels, channels)
        self.to_v = nn.Linear(channels, channels)
        self.to_out = nn.ModuleList([nn.Linear(channels, channels), nn.Identity()])

    def forward(self, x):
        b, c, h, w = x.shape
        t = self.group_norm(x).reshape(b, c, h * w).transpose(1, 2)
        q, k, v = (f(t)[:, None] for f in (self.to_q, self.to_k, self.to_v))
        o = OF.attention(q, k, v, causal=False)[:, 0]
        return x + self.to_out[0](o).transpose(1, 2).reshape(b, c, h, w)


class _Mid(nn.Module):
    def __init__(self, c, groups):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(c, c, None, groups, eps=1e-6), ResnetBlock2D(c, c, None, groups, eps=1e-6)])
        self.attentions = nn.ModuleList([VaeAttention(c, groups)])

    def forward(self, x):
        return self.resnets[1](self.attentions[0](self.resnets[0](x)))


class _EncBlock(nn.Module):
    def __init__(self, cin, cout, layers, down, groups):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if j == 0 else cout, cout, None, groups, eps=1e-6) for j in range(layers)])
        self.downsamplers = nn.ModuleList([Downsample2D(cout, padding=0)]) if down else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        return x if self.downsamplers is None else self.downsamplers[0](x)


class _DecBlock(nn.Module):
    def __init__(self, cin, cout, layers, up, groups):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if j == 0 else cout, cout, None, groups, eps=1e-6) for j in range(layers)])
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if up else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        return x if self.upsamplers is None else self.upsamplers[0](x)


class Encoder(nn.Module):
    def __init__(self, in_channels, latent_channels, ch, layers, groups):
        super().__init__()
        self.conv_in = nn.Conv2d(in_channels, ch[0], 3, padding=1)
        self.down_blocks = nn.ModuleList(
            [_EncBlock(ch[max(i - 1, 0)], ch[i], layers, i < len(ch) - 1, groups) for i in range(len(ch))])
        self.mid_block = _Mid(ch[-1], groups)
        self.conv_norm_out = nn.GroupNorm(groups, ch[-1], eps=1e-6)
        self.conv_out = nn.Conv2d(ch[-1], 2 * latent_channels, 3, padding=1)

    def forward(self, x):
        x = self.conv_in(x)
        for b in self.down_blocks:
            x = b(x)
        return self.conv_out(F.silu(self.conv_norm_out(self.mid_block(x))))


class Decoder(nn.Module):
    def __init__(self, out_channels, latent_channels, ch, layers, groups):
        super().__init__()
        rev = list(reversed(ch))
        self.conv_in = nn.Conv2d(latent_channels, rev[0], 3, padding=1)
        self.mid_block = _Mid(rev[0], groups)
        self.up_blocks = nn.ModuleList(
            [_DecBlock(rev[max(i - 1, 0)], rev[i], layers + 1, i < len(ch) - 1, groups) for i in range(len(ch))])
        self.conv_norm_out = nn.GroupNorm(groups, rev[-1], eps=1e-6)
        self.conv_out = nn.Conv2d(rev[-1], out_channels, 3, padding=1)

    def forward(self, z):
        x = self.mid_block(self.conv_in(z))
        for b in self.up_blocks:
            x = b(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class DiagonalGaussianDistribution:
    def __init__(self, parameters):
        self.mean, logvar = parameters.chunk(2, dim=1)
        self.logvar = logvar.clamp(-30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)

    def sample(self, generator=None):
        noise = torch.randn(self.mean.shape, generator=generator, device=self.mean.device, dtype=self.mean.dtype)
        return self.mean + self.std * noise

    def mode(self):
        return self.mean

    def kl(self):
        return 0.5 * torch.sum(self.mean.pow(2) + self.logvar.exp() - 1.0 - self.logvar, dim=[1, 2, 3])


class _Out:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class AutoencoderKL(nn.Module):
    def __init__(self, in_channels=3, out_channels=3, latent_channels=4, block_out_channels=(128, 256, 512, 512),
                 layers_per_block=2, norm_num_groups=32, scaling_factor=0.18215, **unused):
        super().__init__()
        self.config = dict(in_channels=in_channels, out_channels=out_channels, latent_channels=latent_channels,
                           block_out_channels=list(block_out_channels), layers_per_block=layers_per_block,
                           norm_num_groups=norm_num_groups, scaling_factor=scaling_factor)
        ch = list(block_out_channels)
        self.encoder = Encoder(in_channels, latent_channels, ch, layers_per_block, norm_num_groups)
        self.decoder = Decoder(out_channels, latent_channels, ch, layers_per_block, norm_num_groups)
        self.quant_conv = nn.Conv2d(2 * latent_channels, 2 * latent_channels, 1)
        self.post_quant_conv = nn.Conv2d(latent_channels, latent_channels, 1)

    @property
    def dtype(self):
        return self.quant_conv.weight.dtype

    def encode(self, x):
        return _Out(latent_dist=DiagonalGaussianDistribution(self.quant_conv(self.encoder(x.to(self.dtype)))))

    def decode(self, z):
        return _Out(sample=self.decoder(self.post_quant_conv(z.to(self.dtype))))

    def forward(self, x, sample_posterior=False):
        post = self.encode(x).latent_dist
        return self.decode(post.sample() if sample_posterior else post.mode())
