"""Conditional U-Net of the DALL-E 2 decoder: predicts the noise (and optionally an interpolation coefficient for a
learned variance) of an image given the timestep, a CLIP image embedding and, optionally, the caption's token
encodings and a low-resolution image (cascaded super-resolution stages).

Spec: reference projects/DALLE2/dalle2/models.py:1060-1745 — multi-scale ``CrossEmbedLayer`` stem, ResNet blocks
modulated by a time/image-embedding vector (scale-shift) and cross-attending to conditioning tokens, self-attention
at selected resolutions, classifier-free-guidance dropout of both conditions with learned null embeddings,
``forward_with_cond_scale``.
"""
import math

import torch
import torch.nn.functional as F
from torch import nn

from libai_b200.layers import Linear
from libai_b200.ops import functional as OF

from .prior import SinusoidalPosEmb, prob_mask_like


class ChanLayerNorm(nn.Module):
    def __init__(self, dim, eps=1e-5):
        super().__init__()
        self.eps = eps
        self.g = nn.Parameter(torch.ones(1, dim, 1, 1))

    def forward(self, x):
        var = x.float().var(dim=1, unbiased=False, keepdim=True)
        mean = x.float().mean(dim=1, keepdim=True)
        return ((x.float() - mean) * torch.rsqrt(var + self.eps)).to(x.dtype) * self.g.to(x.dtype)


class CrossEmbedLayer(nn.Module):
    """Stem: parallel convolutions with several kernel sizes, channel-concatenated."""

    def __init__(self, dim_in, kernel_sizes, dim_out=None, stride=1):
        super().__init__()
        dim_out = dim_out or dim_in
        kernel_sizes = sorted(kernel_sizes)
        n = len(kernel_sizes)
        dims = [int(dim_out / (2 ** i)) for i in range(1, n)]
        dims = dims + [dim_out - sum(dims)]
        self.convs = nn.ModuleList([nn.Conv2d(dim_in, d, k, stride=stride, padding=(k - stride) // 2)
                                    for k, d in zip(kernel_sizes, dims)])

    def forward(self, x):
        return torch.cat([c(x) for c in self.convs], dim=1)


class Block(nn.Module):
    def __init__(self, dim, dim_out, groups=8):
        super().__init__()
        self.project = nn.Conv2d(dim, dim_out, 3, padding=1)
        self.norm = nn.GroupNorm(groups, dim_out)

    def forward(self, x, scale_shift=None):
        x = self.norm(self.project(x))
        if scale_shift is not None:
            scale, shift = scale_shift
            x = x * (scale + 1) + shift
        return F.silu(x)


class CrossAttention(nn.Module):
    def __init__(self, dim, *, context_dim=None, dim_head=64, heads=8, norm_context=False):
        super().__init__()
        inner = dim_head * heads
        context_dim = context_dim or dim
        self.heads, self.dim_head = heads, dim_head
        self.norm = nn.LayerNorm(dim)
        self.norm_context = nn.LayerNorm(context_dim) if norm_context else nn.Identity()
        self.null_kv = nn.Parameter(torch.randn(2, dim_head))
        self.to_q = Linear(dim, inner, bias=False)
        self.to_kv = Linear(context_dim, inner * 2, bias=False)
        self.to_out = nn.Sequential(Linear(inner, dim, bias=False), nn.LayerNorm(dim))

    def forward(self, x, context, mask=None):
        b, n, _ = x.shape
        h, d = self.heads, self.dim_head
        q = self.to_q(self.norm(x)).view(b, n, h, d).transpose(1, 2)
        k, v = self.to_kv(self.norm_context(context)).chunk(2, dim=-1)
        k, v = (t.view(b, -1, h, d).transpose(1, 2) for t in (k, v))
        nk, nv = (t.to(k.dtype).view(1, 1, 1, d).expand(b, h, 1, d) for t in self.null_kv.unbind(0))
        k, v = torch.cat([nk, k], dim=2), torch.cat([nv, v], dim=2)
        bias = None
        if mask is not None:
            keep = F.pad(mask, (1, 0), value=True)[:, None, None, :]
            bias = torch.zeros(keep.shape, dtype=q.dtype, device=q.device).masked_fill(~keep, -1e4)
        out = OF.attention(q, k, v, causal=False, bias=bias)
        return self.to_out(out.transpose(1, 2).reshape(b, n, h * d))


class ResnetBlock(nn.Module):
    def __init__(self, dim, dim_out, *, cond_dim=None, time_cond_dim=None, groups=8):
        super().__init__()
        self.time_mlp = nn.Sequential(nn.SiLU(), nn.Linear(time_cond_dim, dim_out * 2)) if time_cond_dim else None
        self.cross_attn = CrossAttention(dim_out, context_dim=cond_dim) if cond_dim else None
        self.block1 = Block(dim, dim_out, groups)
        self.block2 = Block(dim_out, dim_out, groups)
        self.res_conv = nn.Conv2d(dim, dim_out, 1) if dim != dim_out else nn.Identity()

    def forward(self, x, time_emb=None, cond=None, cond_mask=None):
        scale_shift = None
        if self.time_mlp is not None and time_emb is not None:
            scale_shift = self.time_mlp(time_emb).to(x.dtype)[:, :, None, None].chunk(2, dim=1)
        h = self.block1(x)
        if self.cross_attn is not None and cond is not None:
            b, c, hh, ww = h.shape
            t = h.permute(0, 2, 3, 1).reshape(b, hh * ww, c)
            t = self.cross_attn(t, cond, cond_mask) + t
            h = t.reshape(b, hh, ww, c).permute(0, 3, 1, 2)
        h = self.block2(h, scale_shift=scale_shift)
        return h + self.res_conv(x)


class SelfAttention2d(nn.Module):
    """Pre-norm multi-head self-attention over the H·W positions of a feature map."""

    def __init__(self, dim, heads=8, dim_head=64):
        super().__init__()
        self.heads, self.dim_head = heads, dim_head
        self.norm = ChanLayerNorm(dim)
        self.to_qkv = Linear(dim, heads * dim_head * 3, bias=False)
        self.to_out = Linear(heads * dim_head, dim, bias=False)

    def forward(self, x):
        b, c, hh, ww = x.shape
        t = self.norm(x).permute(0, 2, 3, 1).reshape(b, hh * ww, c)
        q, k, v = self.to_qkv(t).view(b, hh * ww, 3, self.heads, self.dim_head).permute(2, 0, 3, 1, 4)
        o = OF.attention(q, k, v, causal=False).transpose(1, 2).reshape(b, hh * ww, -1)
        return x + self.to_out(o).reshape(b, hh, ww, c).permute(0, 3, 1, 2)


class Unet(nn.Module):
    def __init__(self, dim, *, image_embed_dim=None, text_embed_dim=None, cond_dim=None, num_image_tokens=4,
                 num_time_tokens=2, out_dim=None, dim_mults=(1, 2, 4, 8), channels=3, channels_out=None,
                 self_attn=False, attn_dim_head=32, attn_heads=16, lowres_cond=False, sparse_attn=False,
                 cond_on_text_encodings=False, max_text_len=256, cond_on_image_embeds=True, init_dim=None,
                 init_cross_embed_kernel_sizes=(3, 7, 15), num_resnet_blocks=2, resnet_groups=8,
                 memory_efficient=False, **unused):
        super().__init__()
        self._locals = dict(dim=dim, image_embed_dim=image_embed_dim, text_embed_dim=text_embed_dim, cond_dim=cond_dim,
                            dim_mults=tuple(dim_mults), channels=channels, lowres_cond=lowres_cond)
        self.lowres_cond = lowres_cond
        self.channels = channels
        self.channels_out = channels_out or channels
        init_channels = channels * (2 if lowres_cond else 1)
        init_dim = init_dim or dim
        self.init_conv = CrossEmbedLayer(init_channels, init_cross_embed_kernel_sizes, init_dim)
        dims = [init_dim] + [dim * m for m in dim_mults]
        in_out = list(zip(dims[:-1], dims[1:]))
        n_stages = len(in_out)
        cond_dim = cond_dim or dim
        time_cond_dim = dim * 4

        self.to_time_hiddens = nn.Sequential(SinusoidalPosEmb(dim), nn.Linear(dim, time_cond_dim), nn.GELU())
        self.to_time_tokens = nn.Linear(time_cond_dim, cond_dim * num_time_tokens)
        self.to_time_cond = nn.Linear(time_cond_dim, time_cond_dim)
        self.num_time_tokens, self.num_image_tokens, self.cond_dim = num_time_tokens, num_image_tokens, cond_dim

        self.cond_on_image_embeds = cond_on_image_embeds and image_embed_dim is not None
        if self.cond_on_image_embeds:
            self.image_to_tokens = nn.Linear(image_embed_dim, cond_dim * num_image_tokens)
            self.to_image_hiddens = nn.Sequential(nn.Linear(image_embed_dim, time_cond_dim), nn.GELU())
            self.null_image_embed = nn.Parameter(torch.randn(1, num_image_tokens, cond_dim))
            self.null_image_hiddens = nn.Parameter(torch.randn(1, time_cond_dim))
        self.cond_on_text_encodings = cond_on_text_encodings
        self.max_text_len = max_text_len
        if cond_on_text_encodings:
            assert text_embed_dim is not None
            self.text_to_cond = nn.Linear(text_embed_dim, cond_dim)
            self.null_text_embed = nn.Parameter(torch.randn(1, max_text_len, cond_dim))
        self.norm_cond = nn.LayerNorm(cond_dim)

        def per_stage(v):
            return list(v) if hasattr(v, "__iter__") and not isinstance(v, str) else [v] * n_stages

        self_attn, n_blocks = per_stage(self_attn), per_stage(num_resnet_blocks)
        self.downs, self.ups = nn.ModuleList(), nn.ModuleList()
        skip_dims = []
        for i, ((din, dout), attn, nb) in enumerate(zip(in_out, self_attn, n_blocks)):
            last = i == n_stages - 1
            layer_cond = cond_dim if i > 0 else None            # no cross-attention at full resolution
            self.downs.append(nn.ModuleList([
                ResnetBlock(din, dout, time_cond_dim=time_cond_dim, groups=resnet_groups),
                nn.ModuleList([ResnetBlock(dout, dout, cond_dim=layer_cond, time_cond_dim=time_cond_dim, groups=resnet_groups)
                               for _ in range(nb)]),
                SelfAttention2d(dout, attn_heads, attn_dim_head) if attn else nn.Identity(),
                nn.Conv2d(dout, dout, 4, 2, 1) if not last else nn.Identity(),
            ]))
            skip_dims.append(dout)
        mid = dims[-1]
        self.mid_block1 = ResnetBlock(mid, mid, cond_dim=cond_dim, time_cond_dim=time_cond_dim, groups=resnet_groups)
        self.mid_attn = SelfAttention2d(mid, attn_heads, attn_dim_head)
        self.mid_block2 = ResnetBlock(mid, mid, cond_dim=cond_dim, time_cond_dim=time_cond_dim, groups=resnet_groups)
        for i, ((din, dout), attn, nb) in enumerate(zip(reversed(in_out), reversed(self_attn), reversed(n_blocks))):
            last = i == n_stages - 1
            layer_cond = cond_dim if not last else None
            self.ups.append(nn.ModuleList([
                ResnetBlock(dout + skip_dims.pop(), din, cond_dim=layer_cond, time_cond_dim=time_cond_dim, groups=resnet_groups),
                nn.ModuleList([ResnetBlock(din, din, cond_dim=layer_cond, time_cond_dim=time_cond_dim, groups=resnet_groups)
                               for _ in range(nb)]),
                SelfAttention2d(din, attn_heads, attn_dim_head) if attn else nn.Identity(),
                nn.Sequential(nn.Upsample(scale_factor=2, mode="nearest"), nn.Conv2d(din, din, 3, padding=1))
                if not last else nn.Identity(),
            ]))
        self.final_conv = nn.Sequential(ResnetBlock(dims[0], dim, groups=resnet_groups), nn.Conv2d(dim, self.channels_out, 1))

    def cast_model_parameters(self, *, lowres_cond, channels, channels_out, cond_on_image_embeds, cond_on_text_encodings):
        """The decoder may need a differently-shaped copy (e.g. 2× output channels for a learned variance)."""
        same = (lowres_cond == self.lowres_cond and channels == self.channels and channels_out == self.channels_out
                and cond_on_image_embeds == self.cond_on_image_embeds and cond_on_text_encodings == self.cond_on_text_encodings)
        return self, same

    def forward_with_cond_scale(self, *args, cond_scale=1.0, **kwargs):
        logits = self.forward(*args, **kwargs)
        if cond_scale == 1:
            return logits
        null = self.forward(*args, text_cond_drop_prob=1.0, image_cond_drop_prob=1.0, **kwargs)
        return null + (logits - null) * cond_scale

    def forward(self, x, time, *, image_embed=None, lowres_cond_img=None, text_encodings=None, text_mask=None,
                image_cond_drop_prob=0.0, text_cond_drop_prob=0.0):
        b, device = x.shape[0], x.device
        dtype = self.final_conv[1].weight.dtype
        x = x.to(dtype)
        if self.lowres_cond:
            assert lowres_cond_img is not None
            x = torch.cat([x, lowres_cond_img.to(dtype)], dim=1)
        x = self.init_conv(x)

        time_hiddens = self.to_time_hiddens(time).to(dtype)
        time_tokens = self.to_time_tokens(time_hiddens).view(b, self.num_time_tokens, self.cond_dim)
        t = self.to_time_cond(time_hiddens)
        c = time_tokens
        c_mask = torch.ones(b, self.num_time_tokens, dtype=torch.bool, device=device)
        if self.cond_on_image_embeds and image_embed is not None:
            keep = ~prob_mask_like((b,), image_cond_drop_prob, device)
            image_embed = image_embed.to(dtype)
            hiddens = torch.where(keep[:, None], self.to_image_hiddens(image_embed), self.null_image_hiddens.to(dtype))
            t = t + hiddens
            tokens = self.image_to_tokens(image_embed).view(b, self.num_image_tokens, self.cond_dim)
            tokens = torch.where(keep[:, None, None], tokens, self.null_image_embed.to(dtype))
            c = torch.cat([c, tokens], dim=1)
            c_mask = F.pad(c_mask, (0, self.num_image_tokens), value=True)
        if self.cond_on_text_encodings and text_encodings is not None:
            keep = ~prob_mask_like((b,), text_cond_drop_prob, device)
            te = self.text_to_cond(text_encodings.to(dtype))[:, :self.max_text_len]
            tm = (text_mask if text_mask is not None else (text_encodings != 0).any(dim=-1))[:, :self.max_text_len]
            pad = self.max_text_len - te.shape[1]
            if pad > 0:
                te, tm = F.pad(te, (0, 0, 0, pad)), F.pad(tm, (0, pad), value=False)
            te = torch.where((tm & keep[:, None])[..., None], te, self.null_text_embed.to(dtype))
            c = torch.cat([c, te], dim=1)
            c_mask = F.pad(c_mask, (0, te.shape[1]), value=True)
        c = self.norm_cond(c)

        skips = []
        for init_block, blocks, attn, down in self.downs:
            x = init_block(x, t)
            for blk in blocks:
                x = blk(x, t, c, c_mask)
            x = attn(x)
            skips.append(x)
            x = down(x)
        x = self.mid_block1(x, t, c, c_mask)
        x = self.mid_attn(x)
        x = self.mid_block2(x, t, c, c_mask)
        for init_block, blocks, attn, up in self.ups:
            x = init_block(torch.cat([x, skips.pop()], dim=1), t, c, c_mask)
            for blk in blocks:
                x = blk(x, t, c, c_mask)
            x = attn(x)
            x = up(x)
        return self.final_conv[1](self.final_conv[0](x))
