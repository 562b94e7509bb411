"""Diffusion prior of DALL-E 2: a causal transformer that denoises the CLIP *image* embedding conditioned on the
text (tokens ``[text encodings | text embed | time | noised image embed | learned query]`` → prediction read at the
query position).

Spec: reference projects/DALLE2/dalle2/models.py:398-1057 (``RelPosBias``, ``SwiGLU`` feed-forward, ``Attention``
with a null key/value, ``CausalTransformer``, ``DiffusionPriorNetwork`` with classifier-free-guidance dropout and
``forward_with_cond_scale``, ``DiffusionPrior`` training loss and ``sample`` that draws several candidates and keeps
the one closest to the text embedding).

The transformer is tensor-parallel the Megatron way: q/k/v and the first FFN matmul are column-parallel (heads and
FFN channels split over the TP group), output and second FFN matmul row-parallel; a 24×(768, 32 heads) prior on 4
GPUs holds 8 heads per rank.  This is what ``--tensor_parallel 4`` of the reference's inference script selects.
"""
import math

import torch
import torch.nn.functional as F
from torch import nn

from libai_b200.layers import Linear
from libai_b200.utils import distributed as dist

from .diffusion import NoiseScheduler


def l2norm(t):
    return F.normalize(t, dim=-1)


def prob_mask_like(shape, prob, device):
    if prob == 1:
        return torch.ones(shape, device=device, dtype=torch.bool)
    if prob == 0:
        return torch.zeros(shape, device=device, dtype=torch.bool)
    return torch.rand(shape, device=device) < prob


class RelPosBias(nn.Module):
    """T5-style bucketed relative position bias, one scalar per (bucket, head)."""

    def __init__(self, heads=8, num_buckets=32, max_distance=128):
        super().__init__()
        self.num_buckets, self.max_distance = num_buckets, max_distance
        self.relative_attention_bias = nn.Embedding(num_buckets, heads)

    def _bucket(self, rel):
        n = (-rel).clamp(min=0)
        max_exact = self.num_buckets // 2
        large = max_exact + (torch.log(n.float().clamp(min=1) / max_exact) / math.log(self.max_distance / max_exact)
                             * (self.num_buckets - max_exact)).long()
        large = large.clamp(max=self.num_buckets - 1)
        return torch.where(n < max_exact, n, large)

    def forward(self, i, j, device):
        q = torch.arange(j - i, j, device=device)
        k = torch.arange(j, device=device)
        bias = self.relative_attention_bias(self._bucket(k[None, :] - q[:, None]))       # [i, j, heads]
        return bias.permute(2, 0, 1)


class _AllReduceBoth(torch.autograd.Function):
    """Sum over the TP group in forward AND backward (the consumers of the sum differ per rank)."""

    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        x = x.clone()
        torch.distributed.all_reduce(x, group=group)
        return x

    @staticmethod
    def backward(ctx, g):
        g = g.clone()
        torch.distributed.all_reduce(g, group=ctx.group)
        return g, None


class ShardedLayerNorm(nn.Module):
    """LayerNorm over a feature dimension that is split across the tensor-parallel group: the statistics are reduced
    over the group, gamma/beta are the local slices."""

    def __init__(self, full_dim, eps=1e-5):
        super().__init__()
        topo = dist.get_dist_util()
        self.full_dim, self.eps = full_dim, eps
        local = full_dim // topo.tensor_parallel_size
        self.weight = nn.Parameter(torch.ones(local))
        self.bias = nn.Parameter(torch.zeros(local))

    def forward(self, x):
        topo = dist.get_dist_util()
        if topo.tensor_parallel_size == 1:
            return F.layer_norm(x, (x.shape[-1],), self.weight.to(x.dtype), self.bias.to(x.dtype), self.eps)
        xf = x.float()
        stats = torch.stack([xf.sum(-1), (xf * xf).sum(-1)], dim=-1)
        stats = _AllReduceBoth.apply(stats, topo.tp_group)
        mean = stats[..., 0:1] / self.full_dim
        var = stats[..., 1:2] / self.full_dim - mean * mean
        return ((xf - mean) * torch.rsqrt(var + self.eps) * self.weight + self.bias).to(x.dtype)


class SwiGLU(nn.Module):
    def forward(self, x):
        x, gate = x.chunk(2, dim=-1)
        return x * F.silu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim, mult=4, dropout=0.0, post_activation_norm=False):
        super().__init__()
        inner = int(mult * dim)
        self.norm = nn.LayerNorm(dim)
        # value and gate halves are separate column-parallel matmuls so each TP rank holds matching channel slices
        self.w_value = Linear(dim, inner, bias=False, parallel="col")
        self.w_gate = Linear(dim, inner, bias=False, parallel="col")
        self.post_norm = ShardedLayerNorm(inner) if post_activation_norm else None
        self.dropout = nn.Dropout(dropout)
        self.w_out = Linear(inner, dim, bias=False, parallel="row")

    def forward(self, x):
        x = self.norm(x)
        h = self.w_value(x) * F.silu(self.w_gate(x))
        if self.post_norm is not None:
            h = self.post_norm(h)
        return self.w_out(self.dropout(h))


class Attention(nn.Module):
    """Multi-head queries against a single shared key/value head (+ one learned null key/value), cosine-free
    scaled dot product, optional rel-pos bias and causal mask."""

    def __init__(self, dim, *, dim_head=64, heads=8, dropout=0.0, causal=False):
        super().__init__()
        tp = dist.get_dist_util().tensor_parallel_size
        assert heads % tp == 0
        self.scale, self.heads, self.local_heads, self.dim_head, self.causal = dim_head ** -0.5, heads, heads // tp, dim_head, causal
        self.norm = nn.LayerNorm(dim)
        self.dropout = nn.Dropout(dropout)
        self.null_kv = nn.Parameter(torch.randn(2, dim_head))
        self.to_q = Linear(dim, heads * dim_head, bias=False, parallel="col")
        self.to_kv = Linear(dim, 2 * dim_head, bias=False)            # one kv head, replicated
        self.to_out = Linear(heads * dim_head, dim, bias=False, parallel="row")
        self.out_norm = nn.LayerNorm(dim)

    def forward(self, x, mask=None, attn_bias=None):
        b, n, _ = x.shape
        x = self.norm(x)
        q = self.to_q(x).view(b, n, self.local_heads, self.dim_head).transpose(1, 2) * self.scale
        k, v = self.to_kv(x).chunk(2, dim=-1)
        nk, nv = self.null_kv.to(k.dtype).unbind(0)
        k = torch.cat([nk.expand(b, 1, -1), k], dim=1)
        v = torch.cat([nv.expand(b, 1, -1), v], dim=1)
        sim = torch.einsum("bhid,bjd->bhij", q, k)
        if attn_bias is not None:
            sim = sim + attn_bias.to(sim.dtype)
        neg = -torch.finfo(sim.dtype).max
        if mask is not None:
            sim = sim.masked_fill(~F.pad(mask, (1, 0), value=True)[:, None, None, :], neg)
        if self.causal:
            i, j = sim.shape[-2:]
            sim = sim.masked_fill(torch.ones(i, j, dtype=torch.bool, device=x.device).triu(j - i + 1), neg)
        attn = self.dropout(sim.float().softmax(dim=-1).to(sim.dtype))
        out = torch.einsum("bhij,bjd->bhid", attn, v).transpose(1, 2).reshape(b, n, -1)
        return self.out_norm(self.to_out(out))


class CausalTransformer(nn.Module):
    def __init__(self, *, dim, depth, dim_head=64, heads=8, ff_mult=4, norm_out=True, attn_dropout=0.0,
                 ff_dropout=0.0, final_proj=True, normformer=False):
        super().__init__()
        tp = dist.get_dist_util().tensor_parallel_size
        self.heads, self.local_heads = heads, heads // tp
        self.rel_pos_bias = RelPosBias(heads=heads)
        self.layers = nn.ModuleList([
            nn.ModuleList([Attention(dim, causal=True, dim_head=dim_head, heads=heads, dropout=attn_dropout),
                           FeedForward(dim, mult=ff_mult, dropout=ff_dropout, post_activation_norm=normformer)])
            for _ in range(depth)])
        self.norm = nn.LayerNorm(dim) if norm_out else nn.Identity()
        self.project_out = nn.Linear(dim, dim, bias=False) if final_proj else nn.Identity()

    def forward(self, x, mask=None):
        n = x.shape[1]
        bias = self.rel_pos_bias(n, n + 1, device=x.device)                        # +1: the null key
        r = dist.get_dist_util().tp_rank
        bias = bias[r * self.local_heads:(r + 1) * self.local_heads]
        for attn, ff in self.layers:
            x = attn(x, mask=mask, attn_bias=bias) + x
            x = ff(x) + x
        return self.project_out(self.norm(x))


class SinusoidalPosEmb(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.dim = dim

    def forward(self, x):
        half = self.dim // 2
        freq = torch.exp(torch.arange(half, device=x.device, dtype=torch.float32) * -(math.log(10000) / (half - 1)))
        ang = x.float()[:, None] * freq[None]
        return torch.cat([ang.sin(), ang.cos()], dim=-1)


class DiffusionPriorNetwork(nn.Module):
    def __init__(self, dim, num_timesteps=None, num_time_embeds=1, num_image_embeds=1, num_text_embeds=1,
                 max_text_len=256, **kwargs):
        super().__init__()
        self.dim = dim
        self.num_time_embeds, self.num_image_embeds, self.num_text_embeds = num_time_embeds, num_image_embeds, num_text_embeds
        self.to_text_embeds = nn.Linear(dim, dim * num_text_embeds) if num_text_embeds > 1 else nn.Identity()
        self.continuous_embedded_time = num_timesteps is None
        if num_timesteps is not None:
            self.to_time_embeds = nn.Embedding(num_timesteps, dim * num_time_embeds)
        else:
            self.to_time_embeds = nn.Sequential(SinusoidalPosEmb(dim), nn.Linear(dim, dim * 2), nn.SiLU(),
                                                nn.Linear(dim * 2, dim * num_time_embeds))
        self.to_image_embeds = nn.Linear(dim, dim * num_image_embeds) if num_image_embeds > 1 else nn.Identity()
        self.learned_query = nn.Parameter(torch.randn(dim))
        self.causal_transformer = CausalTransformer(dim=dim, **kwargs)
        self.max_text_len = max_text_len
        self.null_text_embed = nn.Parameter(torch.randn(1, max_text_len, dim))

    def forward_with_cond_scale(self, *args, cond_scale=1.0, **kwargs):
        logits = self.forward(*args, **kwargs)
        if cond_scale == 1:
            return logits
        null = self.forward(*args, cond_drop_prob=1.0, **kwargs)
        return null + (logits - null) * cond_scale

    def forward(self, image_embed, diffusion_timesteps, *, text_embed, text_encodings=None, mask=None, cond_drop_prob=0.0):
        b, dim, dtype, device = image_embed.shape[0], self.dim, image_embed.dtype, image_embed.device
        text_embed = self.to_text_embeds(text_embed).view(b, self.num_text_embeds, dim)
        image_embed = self.to_image_embeds(image_embed).view(b, self.num_image_embeds, dim)
        if text_encodings is None:
            text_encodings = torch.empty(b, 0, dim, device=device, dtype=dtype)
        if mask is None:
            mask = (text_encodings != 0).any(dim=-1)
        text_encodings, mask = text_encodings[:, :self.max_text_len], mask[:, :self.max_text_len]
        pad = self.max_text_len - text_encodings.shape[1]
        if pad > 0:
            text_encodings = F.pad(text_encodings, (0, 0, 0, pad))
            mask = F.pad(mask, (0, pad), value=False)
        keep = ~prob_mask_like((b,), cond_drop_prob, device)               # classifier-free guidance dropout
        cond = mask & keep[:, None]
        text_encodings = torch.where(cond[..., None], text_encodings, self.null_text_embed.to(dtype))
        # every token stays visible: dropped / padded text positions now hold the learned null embedding
        time = diffusion_timesteps if not self.continuous_embedded_time else diffusion_timesteps.to(dtype)
        time_embed = self.to_time_embeds(time).view(b, self.num_time_embeds, dim).to(dtype)
        query = self.learned_query.to(dtype).expand(b, 1, dim)
        tokens = torch.cat([text_encodings, text_embed, time_embed, image_embed, query], dim=1)
        return self.causal_transformer(tokens)[:, -1]


class DiffusionPrior(nn.Module):
    def __init__(self, net, *, clip=None, image_embed_dim=None, image_size=None, image_channels=3, timesteps=1000,
                 sample_timesteps=None, cond_drop_prob=0.0, loss_type="l2", predict_x_start=True, beta_schedule="cosine",
                 condition_on_text_encodings=True, sampling_clamp_l2norm=False, training_clamp_l2norm=False,
                 init_image_embed_l2norm=False, image_embed_scale=None, clip_adapter_overrides=None):
        super().__init__()
        self.sample_timesteps = sample_timesteps
        self.noise_scheduler = NoiseScheduler(beta_schedule=beta_schedule, timesteps=timesteps, loss_type=loss_type)
        self.clip = clip
        if clip is not None:
            for p in clip.parameters():
                p.requires_grad_(False)
        self.net = net
        self.image_embed_dim = image_embed_dim or clip.dim_latent
        self.channels = image_channels
        self.cond_drop_prob = cond_drop_prob
        self.can_classifier_guidance = cond_drop_prob > 0.0
        self.condition_on_text_encodings = condition_on_text_encodings
        self.predict_x_start = predict_x_start
        self.image_embed_scale = image_embed_scale or self.image_embed_dim ** 0.5
        self.sampling_clamp_l2norm, self.training_clamp_l2norm = sampling_clamp_l2norm, training_clamp_l2norm
        self.init_image_embed_l2norm = init_image_embed_l2norm

    def l2norm_clamp_embed(self, e):
        return l2norm(e) * self.image_embed_scale

    def p_mean_variance(self, x, t, text_cond, clip_denoised=False, cond_scale=1.0):
        pred = self.net.forward_with_cond_scale(x, t, cond_scale=cond_scale, **text_cond)
        x0 = pred if self.predict_x_start else self.noise_scheduler.predict_start_from_noise(x, t, pred)
        if clip_denoised and not self.predict_x_start:
            x0 = x0.clamp(-1.0, 1.0)
        if self.predict_x_start and self.sampling_clamp_l2norm:
            x0 = self.l2norm_clamp_embed(x0)
        return self.noise_scheduler.q_posterior(x0, x, t)

    @torch.no_grad()
    def p_sample(self, x, t, text_cond=None, clip_denoised=True, cond_scale=1.0):
        mean, _, log_var = self.p_mean_variance(x, t, text_cond, clip_denoised, cond_scale)
        nonzero = (t != 0).float().view(-1, *((1,) * (x.dim() - 1)))
        return mean + nonzero * (0.5 * log_var).exp() * torch.randn_like(x)

    @torch.no_grad()
    def p_sample_loop(self, shape, text_cond, cond_scale=1.0):
        device = self.noise_scheduler.betas.device
        x = torch.randn(shape, device=device)
        if self.init_image_embed_l2norm:
            x = l2norm(x) * self.image_embed_scale
        steps = range(self.noise_scheduler.num_timesteps - 1, -1, -1)
        for i in steps:
            t = torch.full((shape[0],), i, device=device, dtype=torch.long)
            x = self.p_sample(x, t, text_cond=text_cond, cond_scale=cond_scale)
        return x

    def p_losses(self, image_embed, times, text_cond, noise=None):
        noise = torch.randn_like(image_embed) if noise is None else noise
        noisy = self.noise_scheduler.q_sample(image_embed, times, noise)
        pred = self.net(noisy, times, cond_drop_prob=self.cond_drop_prob, **text_cond)
        if self.predict_x_start and self.training_clamp_l2norm:
            pred = self.l2norm_clamp_embed(pred)
        return self.noise_scheduler.loss_fn(pred, image_embed if self.predict_x_start else noise)

    @torch.no_grad()
    def sample(self, text, num_samples_per_batch=2, cond_scale=1.0):
        """Draw ``num_samples_per_batch`` image embeddings per caption and keep the most text-similar one."""
        n = num_samples_per_batch
        text = text.repeat_interleave(n, dim=0)
        b = text.shape[0]
        text_embed, text_encodings, text_mask = self.clip.embed_text(text)
        cond = dict(text_embed=text_embed)
        if self.condition_on_text_encodings:
            cond.update(text_encodings=text_encodings, mask=text_mask)
        embeds = self.p_sample_loop((b, self.image_embed_dim), text_cond=cond, cond_scale=cond_scale) / self.image_embed_scale
        text_embed = text_embed.view(b // n, n, -1)
        embeds = embeds.view(b // n, n, -1)
        sims = torch.einsum("bnd,bnd->bn", l2norm(text_embed.float()), l2norm(embeds.float()))
        top = sims.argmax(dim=1)
        return embeds[torch.arange(b // n, device=embeds.device), top]

    def forward(self, text=None, image=None, text_embed=None, image_embed=None, text_encodings=None, text_mask=None):
        assert (text is None) != (text_embed is None) and (image is None) != (image_embed is None)
        if image is not None:
            image_embed, _ = self.clip.embed_image(image)
        if text is not None:
            text_embed, text_encodings, text_mask = self.clip.embed_text(text)
        cond = dict(text_embed=text_embed)
        if self.condition_on_text_encodings:
            cond.update(text_encodings=text_encodings, mask=text_mask)
        times = self.noise_scheduler.sample_random_times(image_embed.shape[0], image_embed.device)
        return self.p_losses(image_embed * self.image_embed_scale, times, cond)
