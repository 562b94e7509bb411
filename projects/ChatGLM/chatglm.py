"""ChatGLM2/3-6B: RMSNorm, half-dimension interleaved rotary, multi-query attention, SwiGLU, optional P-tuning prefix.

Spec: reference projects/ChatGLM/chatglm.py — ``PrefixEncoder`` (:42-71), ``RotaryEmbedding`` with a
``[seq, d/4, 2]`` cos/sin cache applied to the first half of every head in interleaved pairs (:74-115 and
``apply_rotary_pos_emb``), ``SelfAttention`` with ``multi_query_group_num`` KV heads in one fused
``[a·d | g·d | g·d]`` projection (:228-370), ``MLP`` with a fused ``[2·ffn]`` up-projection and
``silu(x₀)·x₁`` (:373-416), ``GLMBlock`` / ``GLMTransformer`` (:419-577), ``ChatGLMModel`` (:668-785),
``ChatGLMForConditionalGeneration`` with generation hooks and ``chat`` (:788-990).  Parameter names follow the
original checkpoint (``transformer.encoder.layers.N.self_attention.query_key_value`` …) so the HF loader is 1:1.

Parallelism: data / pipeline parallel (the fused multi-query projection is not split over heads here).
"""
from __future__ import annotations

import math
from typing import List, Optional

import torch
from torch import nn

from libai_b200.inference.generator.generation_utils import Generator
from libai_b200.layers import LayerNorm, Linear, RMSLayerNorm
from libai_b200.layers._param import create_parameter
from libai_b200.models.utils.weight_init import init_method_normal
from libai_b200.ops import functional as OF


class PrefixEncoder(nn.Module):
    """P-tuning v2: ``[pre_seq_len]`` virtual tokens → per-layer key/value prefixes."""

    def __init__(self, cfg):
        super().__init__()
        self.prefix_projection = cfg.prefix_projection
        kv_size = cfg.num_layers * cfg.kv_channels * cfg.multi_query_group_num * 2
        if self.prefix_projection:
            self.embedding = nn.Embedding(cfg.pre_seq_len, kv_size)
            self.trans = nn.Sequential(nn.Linear(kv_size, cfg.hidden_size), nn.Tanh(), nn.Linear(cfg.hidden_size, kv_size))
        else:
            self.embedding = nn.Embedding(cfg.pre_seq_len, kv_size)

    def forward(self, prefix):
        tokens = self.embedding(prefix)
        return self.trans(tokens) if self.prefix_projection else tokens


def rope_cache(seq_len: int, rotary_dim: int, base: float = 10000.0, device=None):
    """``[seq, rotary_dim/2, 2]`` (cos, sin) table; ``rotary_dim`` = half of the head dimension."""
    theta = 1.0 / (base ** (torch.arange(0, rotary_dim, 2, dtype=torch.float32, device=device) / rotary_dim))
    idx = torch.outer(torch.arange(seq_len, dtype=torch.float32, device=device), theta)
    return torch.stack([idx.cos(), idx.sin()], dim=-1)


def apply_rotary_pos_emb(x: torch.Tensor, cache: torch.Tensor) -> torch.Tensor:
    """x ``[b, s, heads, d]``; rotates interleaved pairs of the first ``2·cache.shape[-2]`` channels."""
    b, s, h, d = x.shape
    rot = cache.shape[-2] * 2
    xr, xp = x[..., :rot], x[..., rot:]
    xr = xr.float().reshape(b, s, h, rot // 2, 2)
    c = cache[:s].view(1, s, 1, rot // 2, 2)
    out = torch.stack([xr[..., 0] * c[..., 0] - xr[..., 1] * c[..., 1], xr[..., 1] * c[..., 0] + xr[..., 0] * c[..., 1]], -1)
    return torch.cat([out.flatten(3).to(x.dtype), xp], dim=-1)


class RotaryEmbedding(nn.Module):
    """``rope(max_seq_len)`` → the ``[seq, rotary_dim/2, 2]`` (cos, sin) cache consumed by :func:`apply_rotary_pos_emb`
    (reference projects/ChatGLM/chatglm.py RotaryEmbedding)."""

    def __init__(self, dim, original_impl=False, device=None, dtype=None, base: float = 10000.0):
        super().__init__()
        self.dim, self.original_impl, self.base = dim, original_impl, base

    def forward(self, max_seq_len, offset=0, device=None):
        return rope_cache(max_seq_len, self.dim, self.base, device=device)


class CoreAttention(nn.Module):
    """Scaled-dot-product core on ``[b, heads, s, d]`` tensors → ``[b, s, heads·d]``: the flash kernel when no explicit
    mask is given (causal for square scores), the masked reference math otherwise."""

    def __init__(self, cfg=None, layer_number=1):
        super().__init__()
        self.layer_number = max(1, layer_number)
        self.attention_dropout = float(cfg.attention_dropout) if cfg is not None else 0.0

    def forward(self, query_layer, key_layer, value_layer, attention_mask=None):
        b, a, s, d = query_layer.shape
        causal = attention_mask is None and s == key_layer.shape[2]
        ctx = OF.attention(query_layer, key_layer, value_layer, causal=causal, scale=1.0 / math.sqrt(d),
                           mask=attention_mask, dropout_p=self.attention_dropout, training=self.training)
        return ctx.transpose(1, 2).reshape(b, s, a * d)


class SelfAttention(nn.Module):
    def __init__(self, cfg, layer_number):
        super().__init__()
        self.layer_number = max(1, layer_number)
        self.heads, self.d = cfg.num_attention_heads, cfg.kv_channels
        self.groups = cfg.multi_query_group_num if cfg.multi_query_attention else cfg.num_attention_heads
        proj = self.heads * self.d
        init = init_method_normal(0.02)
        self.query_key_value = Linear(cfg.hidden_size, proj + 2 * self.groups * self.d,
                                      bias=cfg.add_bias_linear or cfg.add_qkv_bias, parallel="data", init_method=init,
                                      layer_idx=layer_number - 1)
        self.dense = Linear(proj, cfg.hidden_size, bias=cfg.add_bias_linear, parallel="data", init_method=init,
                            layer_idx=layer_number - 1)
        self.attention_dropout = cfg.attention_dropout
        self.core_attention = CoreAttention(cfg, layer_number)

    def forward(self, hidden, attention_mask, rotary_pos_emb, kv_cache=None, use_cache=True):
        b, s, _ = hidden.shape
        a, g, d = self.heads, self.groups, self.d
        mixed = self.query_key_value(hidden)
        q, k, v = mixed.split([a * d, g * d, g * d], dim=-1)
        q, k, v = q.view(b, s, a, d), k.view(b, s, g, d), v.view(b, s, g, d)
        if rotary_pos_emb is not None:
            q, k = apply_rotary_pos_emb(q, rotary_pos_emb), apply_rotary_pos_emb(k, rotary_pos_emb)
        q, k, v = q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2)  # [b, heads, s, d]
        if kv_cache is not None:
            k = torch.cat((kv_cache[0].type_as(k), k), dim=2)
            v = torch.cat((kv_cache[1].type_as(v), v), dim=2)
        new_cache = (k, v) if use_cache else None
        if g != a:  # multi-query: every group of a/g query heads shares one KV head
            k = k.repeat_interleave(a // g, dim=1)
            v = v.repeat_interleave(a // g, dim=1)
        return self.dense(self.core_attention(q, k, v, attention_mask)), new_cache


class MLP(nn.Module):
    def __init__(self, cfg, layer_idx):
        super().__init__()
        init = init_method_normal(0.02)
        self.dense_h_to_4h = Linear(cfg.hidden_size, cfg.ffn_hidden_size * 2, bias=cfg.add_bias_linear, parallel="data",
                                    init_method=init, layer_idx=layer_idx)
        self.dense_4h_to_h = Linear(cfg.ffn_hidden_size, cfg.hidden_size, bias=cfg.add_bias_linear, parallel="data",
                                    init_method=init, layer_idx=layer_idx)

    def forward(self, hidden):
        gate, up = self.dense_h_to_4h(hidden).chunk(2, dim=-1)
        return self.dense_4h_to_h(OF.swiglu(gate.contiguous(), up.contiguous()))


class GLMBlock(nn.Module):
    def __init__(self, cfg, layer_number):
        super().__init__()
        self.layer_number, self.layer_idx = layer_number, layer_number - 1
        self.apply_residual_connection_post_layernorm = cfg.apply_residual_connection_post_layernorm
        norm = RMSLayerNorm if cfg.rmsnorm else LayerNorm
        self.input_layernorm = norm(cfg.hidden_size, eps=cfg.layernorm_epsilon, layer_idx=self.layer_idx)
        self.self_attention = SelfAttention(cfg, layer_number)
        self.hidden_dropout = cfg.hidden_dropout
        self.post_attention_layernorm = norm(cfg.hidden_size, eps=cfg.layernorm_epsilon, layer_idx=self.layer_idx)
        self.mlp = MLP(cfg, self.layer_idx)

    def forward(self, hidden, attention_mask, rotary_pos_emb, kv_cache=None, use_cache=True):
        ln = self.input_layernorm(hidden)
        attn, kv_cache = self.self_attention(ln, attention_mask, rotary_pos_emb, kv_cache=kv_cache, use_cache=use_cache)
        residual = ln if self.apply_residual_connection_post_layernorm else hidden
        x = residual + torch.nn.functional.dropout(attn, p=self.hidden_dropout, training=self.training)
        ln2 = self.post_attention_layernorm(x)
        residual = ln2 if self.apply_residual_connection_post_layernorm else x
        out = residual + torch.nn.functional.dropout(self.mlp(ln2), p=self.hidden_dropout, training=self.training)
        return out, kv_cache


class GLMTransformer(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.num_layers, self.post_layer_norm = cfg.num_layers, cfg.post_layer_norm
        self.layers = nn.ModuleList([GLMBlock(cfg, i + 1) for i in range(cfg.num_layers)])
        if self.post_layer_norm:
            norm = RMSLayerNorm if cfg.rmsnorm else LayerNorm
            self.final_layernorm = norm(cfg.hidden_size, eps=cfg.layernorm_epsilon, layer_idx=-1)

    def forward(self, hidden, attention_mask, rotary_pos_emb, kv_caches=None, use_cache=True):
        kv_caches = kv_caches or [None] * self.num_layers
        presents = []
        for layer, cache in zip(self.layers, kv_caches):
            hidden, present = layer(hidden, attention_mask, rotary_pos_emb, kv_cache=cache, use_cache=use_cache)
            presents.append(present)
        if self.post_layer_norm:
            hidden = self.final_layernorm(hidden)
        return hidden, (presents if use_cache else None)


class EmbeddingLayer(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.word_embeddings = nn.Embedding(cfg.padded_vocab_size, cfg.hidden_size)

    def forward(self, input_ids):
        return self.word_embeddings(input_ids)


class ChatGLMPreTrainedModel:
    """Helpers shared by the ChatGLM model classes (reference projects/ChatGLM/chatglm.py ChatGLMPreTrainedModel:
    ``_init_weights`` / ``get_masks`` / ``get_position_ids``); a mixin — parameters are initialised by the layer
    constructors, so ``_init_weights`` has nothing left to do."""

    def _init_weights(self, module: nn.Module):
        return

    def get_position_ids(self, input_ids):
        b, s = input_ids.shape
        return torch.arange(s, dtype=torch.long, device=input_ids.device).unsqueeze(0).repeat(b, 1)


class ChatGLMModel(nn.Module, ChatGLMPreTrainedModel):
    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        self.embedding = EmbeddingLayer(cfg)
        self.num_layers, self.groups, self.kv_channels = cfg.num_layers, cfg.multi_query_group_num, cfg.kv_channels
        self.seq_length = cfg.seq_length
        self.rotary_dim = cfg.kv_channels // 2
        self.encoder = GLMTransformer(cfg)
        self.output_layer = Linear(cfg.hidden_size, cfg.padded_vocab_size, bias=False, parallel="data",
                                   init_method=init_method_normal(0.02), layer_idx=-1)
        self.pre_seq_len = cfg.pre_seq_len
        if self.pre_seq_len is not None:
            for p in self.parameters():
                p.requires_grad = False
            self.register_buffer("prefix_tokens", torch.arange(self.pre_seq_len).long(), persistent=False)
            self.prefix_encoder = PrefixEncoder(cfg)
            self.dropout = nn.Dropout(0.1)
        self._rope = None

    def get_input_embeddings(self):
        return self.embedding.word_embeddings

    def get_prompt(self, batch_size, device, dtype):
        tokens = self.prefix_tokens.unsqueeze(0).expand(batch_size, -1).to(device)
        kv = self.prefix_encoder(tokens).to(dtype)
        kv = kv.view(batch_size, self.pre_seq_len, self.num_layers * 2, self.groups, self.kv_channels)
        kv = self.dropout(kv).permute(2, 0, 3, 1, 4)  # [layers*2, b, groups, pre, d]
        return [(kv[2 * i], kv[2 * i + 1]) for i in range(self.num_layers)]

    def get_masks(self, input_ids, past_length, padding_mask=None):
        """Boolean visibility ``[b, 1, q, k]`` (None = plain causal without cache → kernel handles it)."""
        b, q = input_ids.shape
        if past_length == 0 and (padding_mask is None or bool(padding_mask.all())):
            return None
        k = past_length + q
        mask = torch.ones(k, k, dtype=torch.bool, device=input_ids.device).tril()[k - q :][None, None].expand(b, 1, q, k)
        if padding_mask is not None:
            mask = mask & padding_mask.bool()[:, None, None, :k]
        return mask

    def forward(self, input_ids, position_ids=None, attention_mask=None, past_key_values=None, use_cache=False):
        b, q = input_ids.shape
        hidden = self.embedding(input_ids)
        if self.pre_seq_len is not None and past_key_values is None:
            past_key_values = self.get_prompt(b, input_ids.device, hidden.dtype)
        past_len = 0 if not past_key_values or past_key_values[0] is None else past_key_values[0][0].shape[2]
        if self._rope is None or self._rope.device != input_ids.device or self._rope.shape[0] < self.seq_length:
            self._rope = rope_cache(self.seq_length, self.rotary_dim, device=input_ids.device)
        rope_offset = past_len - (self.pre_seq_len or 0) if self.pre_seq_len is not None else past_len
        if position_ids is not None:
            rope = self._rope[position_ids[0]]
        else:
            rope = self._rope[max(rope_offset, 0) : max(rope_offset, 0) + q]
        mask = self.get_masks(input_ids, past_len, attention_mask)
        hidden, presents = self.encoder(hidden, mask, rope, kv_caches=past_key_values, use_cache=use_cache)
        return hidden, presents


class ChatGLMForConditionalGeneration(nn.Module, ChatGLMPreTrainedModel, Generator):
    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        self.transformer = ChatGLMModel(cfg)
        self.past_key_values: List = [None] * cfg.num_layers
        if cfg.get("lora_enable", False):
            from projects.ChatGLM.lora.lora_model import LoraModel

            self.transformer = LoraModel(self.transformer, cfg.lora_cfg, "default")

    def forward(self, input_ids, position_ids=None, attention_mask=None, labels=None, use_cache=False):
        past = self.past_key_values if use_cache and self.past_key_values[0] is not None else None
        hidden, presents = self.transformer(input_ids, position_ids, attention_mask, past, use_cache)
        if use_cache:
            self.set_cache(presents)
        model = self.transformer.model if hasattr(self.transformer, "model") else self.transformer
        logits = model.output_layer(hidden)
        if labels is not None:
            shift_logits = logits[:, :-1].float().reshape(-1, logits.shape[-1])
            shift_labels = labels[:, 1:].reshape(-1)
            loss = torch.nn.functional.cross_entropy(shift_logits, shift_labels, ignore_index=-100)
            return {"loss": loss}
        return {"logits": logits}

    def set_cache(self, past_key_values):
        self.past_key_values = [None] * self.cfg.num_layers if past_key_values is None else list(past_key_values)

    def prepare_inputs_for_generation(self, input_ids, past=None, attention_mask=None, use_cache=None, **kwargs):
        if past is not None and use_cache:
            input_ids = input_ids[:, -1:]
        return {"input_ids": input_ids, "attention_mask": attention_mask, "use_cache": bool(use_cache)}

    def process_response(self, output, history):
        content = ""
        history = list(history)
        for response in output.split("<|assistant|>"):
            if "\n" in response:
                metadata, content = response.split("\n", maxsplit=1)
            else:
                metadata, content = "", response
            if not metadata.strip():
                content = content.strip().replace("[[训练时间]]", "2023年")
                history.append({"role": "assistant", "metadata": metadata, "content": content})
            else:
                history.append({"role": "assistant", "metadata": metadata, "content": content})
        return content, history

    @torch.no_grad()
    def chat(self, tokenizer, query, history=None, role="user", max_length=8192, num_beams=1, do_sample=True, top_p=0.8,
             temperature=0.8, **kwargs):
        history = history or []
        inputs = tokenizer.build_chat_input(query, history=history, role=role)
        ids = inputs["input_ids"].to(next(self.parameters()).device)
        out = self.generate(ids, max_length=max_length, num_beams=num_beams, do_sample=do_sample, top_p=top_p,
                            temperature=temperature, eos_token_id=tokenizer.eos_token_id, **kwargs)
        response = tokenizer.decode(out[0, ids.shape[1] : -1].tolist())
        history.append({"role": role, "content": query})
        return self.process_response(response, history)

    @staticmethod
    def set_activation_checkpoint(model):
        model.activation_checkpoint = True
        return model
