"""BLOOM: ALiBi causal decoder (LayerNorm, GELU, tied embeddings).

Spec: reference projects/BLOOM/modeling/{bloom_model.py,attention.py,mask.py,mlp.py,transformers.py} —
``build_alibi_tensor`` (mask.py:65-101: per-head slopes ``2^(-8i/n)`` with the closest-power-of-two rule, bias =
slope × key position counted over the attention mask), ``BloomAttention`` (fused per-head interleaved qkv,
``baddbmm(alibi, q, kᵀ, beta=1, alpha=1/sqrt(d))``), ``BloomModel`` (:29-255) with embedding LayerNorm and a
final ``ln_f``, ``BloomForCausalLM`` (:258-420) with the generation hooks.  Names follow the HF checkpoint layout
(``word_embeddings``, ``h.N.self_attention.query_key_value`` …) so the loader is a prefix strip.
"""
from __future__ import annotations

import math

import torch
from torch import nn

from libai_b200.config import configurable
from libai_b200.inference.generator.generation_utils import Generator
from libai_b200.layers import LayerNorm, Linear, LMLogits, ParallelCrossEntropyLoss, VocabEmbedding
from libai_b200.models.utils.weight_init import init_method_normal
from libai_b200.ops import functional as OF
from libai_b200.parallel import mappings
from libai_b200.utils import distributed as dutil


def alibi_slopes(num_heads: int) -> torch.Tensor:
    closest = 2 ** math.floor(math.log2(num_heads))
    base = 2.0 ** (-(2.0 ** -(math.log2(closest) - 3)))
    slopes = torch.pow(torch.tensor(base), torch.arange(1, closest + 1, dtype=torch.float32))
    if closest != num_heads:
        extra_base = 2.0 ** (-(2.0 ** -(math.log2(2 * closest) - 3)))
        n_extra = min(closest, num_heads - closest)
        slopes = torch.cat([slopes, torch.pow(torch.tensor(extra_base), torch.arange(1, 1 + 2 * n_extra, 2, dtype=torch.float32))])
    return slopes


def build_alibi_tensor(attention_mask: torch.Tensor, num_heads: int, dtype=torch.float32) -> torch.Tensor:
    """``[b, heads, 1, k]`` additive bias; key positions are counted over the non-padded tokens."""
    pos = ((attention_mask.long().cumsum(dim=-1) - 1) * attention_mask.long())[:, None, None, :]
    return (alibi_slopes(num_heads).to(attention_mask.device)[None, :, None, None] * pos).to(dtype)


class BloomAttention(nn.Module):
    def __init__(self, hidden_size, n_head, hidden_dropout, attention_dropout, init_method, layer_idx):
        super().__init__()
        topo = dutil.get_dist_util()
        self.num_heads, self.head_dim = n_head, hidden_size // n_head
        self.local_heads = n_head // topo.tensor_parallel_size
        self.query_key_value = Linear(hidden_size, 3 * hidden_size, bias=True, parallel="col", init_method=init_method, layer_idx=layer_idx)
        self.dense = Linear(hidden_size, hidden_size, bias=True, parallel="row", init_method=init_method, layer_idx=layer_idx)
        self.attention_dropout, self.hidden_dropout = attention_dropout, nn.Dropout(hidden_dropout)

    def forward(self, hidden, residual, alibi, mask, layer_past=None, use_cache=False):
        b = hidden.shape[0]
        a, d = self.local_heads, self.head_dim
        qkv = self.query_key_value(hidden).view(b, -1, a, 3 * d).permute(0, 2, 1, 3)
        q, k, v = qkv[..., :d], qkv[..., d : 2 * d], qkv[..., 2 * d :]
        if layer_past is not None:
            k = torch.cat((layer_past[0].type_as(k), k), dim=2)
            v = torch.cat((layer_past[1].type_as(v), v), dim=2)
        present = (k, v) if use_cache else None
        topo = dutil.get_dist_util()
        if alibi.dim() == 1:
            # unpadded batch without a KV cache: ALiBi slopes + causal mask are evaluated inside the flash kernel
            slopes = alibi[topo.tp_rank * a : (topo.tp_rank + 1) * a] if alibi.shape[0] != a else alibi
            ctx = OF.attention(q, k, v, causal=True, scale=1.0 / math.sqrt(d), alibi_slopes=slopes.contiguous(),
                               dropout_p=self.attention_dropout, training=self.training)
        else:
            bias = alibi[:, topo.tp_rank * a : (topo.tp_rank + 1) * a] if alibi.shape[1] != a else alibi
            ctx = OF.attention(q, k, v, causal=False, scale=1.0 / math.sqrt(d), mask=mask, bias=bias,
                               dropout_p=self.attention_dropout, training=self.training)
        out = self.dense(ctx.transpose(1, 2).reshape(b, -1, a * d))
        return residual + self.hidden_dropout(out), present


class BloomMLP(nn.Module):
    def __init__(self, hidden_size, hidden_dropout, init_method, layer_idx):
        super().__init__()
        self.dense_h_to_4h = Linear(hidden_size, 4 * hidden_size, parallel="col", init_method=init_method, layer_idx=layer_idx)
        self.dense_4h_to_h = Linear(4 * hidden_size, hidden_size, parallel="row", init_method=init_method, layer_idx=layer_idx)
        self.dropout = nn.Dropout(hidden_dropout)

    def forward(self, hidden, residual):
        return residual + self.dropout(self.dense_4h_to_h(self.dense_h_to_4h(hidden, act="gelu_tanh")))


class BloomBlock(nn.Module):
    def __init__(self, hidden_size, n_head, layer_norm_epsilon, hidden_dropout, attention_dropout,
                 apply_residual_connection_post_layernorm, init_method, layer_idx):
        super().__init__()
        self.layer_idx = layer_idx
        self.input_layernorm = LayerNorm(hidden_size, eps=layer_norm_epsilon, layer_idx=layer_idx)
        self.self_attention = BloomAttention(hidden_size, n_head, hidden_dropout, attention_dropout, init_method, layer_idx)
        self.post_attention_layernorm = LayerNorm(hidden_size, eps=layer_norm_epsilon, layer_idx=layer_idx)
        self.mlp = BloomMLP(hidden_size, hidden_dropout, init_method, layer_idx)
        self.post_ln_residual = apply_residual_connection_post_layernorm

    def forward(self, hidden, alibi, mask, layer_past=None, use_cache=False):
        ln = self.input_layernorm(hidden)
        attn, present = self.self_attention(ln, ln if self.post_ln_residual else hidden, alibi, mask, layer_past, use_cache)
        ln2 = self.post_attention_layernorm(attn)
        return self.mlp(ln2, ln2 if self.post_ln_residual else attn), present


class BloomModel(nn.Module):
    @configurable
    def __init__(self, vocab_size, hidden_size, hidden_layers, n_head, padding_idx=None, pretraining_tp=1,
                 slow_but_exact=False, initializer_range=0.02, apply_residual_connection_post_layernorm=False,
                 hidden_dropout=0.0, attention_dropout=0.0, amp_enabled=False, layer_norm_epsilon=1e-12, cfg=None):
        super().__init__()
        self.cfg, self.n_head, self.hidden_layers = cfg, n_head, hidden_layers
        init_method = init_method_normal(initializer_range)
        self.word_embeddings = VocabEmbedding(vocab_size, hidden_size, padding_idx=padding_idx, init_method=init_method,
                                              amp_enabled=amp_enabled)
        self.word_embeddings_layernorm = LayerNorm(hidden_size, eps=layer_norm_epsilon, layer_idx=0)
        self.h = nn.ModuleList([
            BloomBlock(hidden_size, n_head, layer_norm_epsilon, hidden_dropout, attention_dropout,
                       apply_residual_connection_post_layernorm, init_method, layer_idx=i)
            for i in range(hidden_layers)])
        self.ln_f = LayerNorm(hidden_size, eps=layer_norm_epsilon, layer_idx=-1)

    @classmethod
    def from_config(cls, cfg):
        keys = ("vocab_size hidden_size hidden_layers n_head padding_idx pretraining_tp slow_but_exact initializer_range "
                "apply_residual_connection_post_layernorm hidden_dropout attention_dropout amp_enabled layer_norm_epsilon").split()
        out = {k: cfg[k] for k in keys if k in cfg}
        out["cfg"] = cfg
        return out

    def forward(self, input_ids, attention_mask=None, past_key_values=None, use_cache=False):
        b, q = input_ids.shape
        past_key_values = past_key_values or [None] * len(self.h)
        past_len = 0 if past_key_values[0] is None else past_key_values[0][0].shape[2]
        k = past_len + q
        hidden = self.word_embeddings_layernorm(self.word_embeddings(input_ids))
        if attention_mask is None and past_len == 0 and not use_cache and input_ids.is_cuda:
            # no padding, no cache: hand the blocks the per-head slopes only (1-D `alibi` selects the in-kernel path)
            alibi, mask = alibi_slopes(self.n_head).to(input_ids.device), None
        else:
            if attention_mask is None:
                attention_mask = torch.ones(b, k, dtype=torch.long, device=input_ids.device)
            alibi = build_alibi_tensor(attention_mask, self.n_head, torch.float32)
            causal = torch.ones(k, k, dtype=torch.bool, device=input_ids.device).tril()[k - q :]
            mask = causal[None, None] & attention_mask.bool()[:, None, None, :]
        presents = []
        for block, past in zip(self.h, past_key_values):
            hidden, present = block(hidden, alibi, mask, past, use_cache)
            presents.append(present)
        return self.ln_f(hidden), (presents if use_cache else None)


class BloomForCausalLM(nn.Module, Generator):
    @configurable
    def __init__(self, cfg=None, **kwargs):
        super().__init__()
        self.cfg = cfg
        self.transformer = BloomModel(cfg=cfg, **kwargs) if cfg is None else BloomModel(cfg)
        self.lm_head = LMLogits(self.transformer.word_embeddings.num_embeddings, bias=False)
        self.loss = ParallelCrossEntropyLoss()
        self.past_key_values = [None] * self.transformer.hidden_layers

    @classmethod
    def from_config(cls, cfg):
        return {"cfg": cfg}

    def forward(self, input_ids, attention_mask=None, labels=None, use_cache=False):
        hidden, presents = self.transformer(input_ids, attention_mask,
                                            self.past_key_values if use_cache else None, use_cache)
        if use_cache:
            self.set_cache(presents)
        logits = self.lm_head(hidden, self.transformer.word_embeddings.weight)
        if labels is not None:
            shift = self.loss(logits[:, :-1].contiguous(), labels[:, 1:].contiguous())
            return {"loss": shift.mean()}
        if dutil.get_dist_util().tensor_parallel_size > 1:
            logits = mappings.gather_from_tp(logits)
        return {"logits": logits}

    def set_cache(self, past_key_values):
        self.past_key_values = [None] * self.transformer.hidden_layers if past_key_values is None else list(past_key_values)

    def prepare_inputs_for_generation(self, input_ids, past=None, attention_mask=None, use_cache=None, **kwargs):
        if past is not None and use_cache:
            input_ids = input_ids[:, -1:]
        return {"input_ids": input_ids, "attention_mask": attention_mask, "use_cache": bool(use_cache)}
