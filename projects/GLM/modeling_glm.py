"""GLM: autoregressive blank-infilling transformer (2-D positions, prefix-LM masks, hidden-state memories).

Spec: reference projects/GLM/modeling_glm.py — ``Transformer`` (:30-80, pre-LN GPT-style blocks whose attention
reads ``[memory; hidden]`` for keys/values), ``GLMModel`` (:83-260: embeddings with position + block-position
tables, ``build_mask_matrix`` from a scalar / per-sample separator = bidirectional context + causal generation part,
``update_mems`` keeping the inputs of every layer as the generation cache), ``GLMLoss``, ``GLMForMultipleChoice``
(:273-309), ``GLMForConditionalGeneration`` (:312-470) and projects/GLM/layers/*.
"""
from __future__ import annotations

import math

import torch
from torch import nn

from libai_b200.config import configurable
from libai_b200.inference.generator.generation_utils import Generator
from libai_b200.layers import Embedding, LayerNorm, Linear, LMLogits, MLP, VocabEmbedding
from libai_b200.models.utils.weight_init import init_method_normal, scaled_init_method_normal
from libai_b200.ops import functional as OF
from libai_b200.parallel import mappings
from libai_b200.utils import distributed as dutil


class GLMEmbedding(nn.Module):
    def __init__(self, vocab_size, hidden_size, max_seq_length, padding_idx=None, init_method=None,
                 embedding_dropout_prob=0.0, amp_enabled=False, block_position_encoding=False):
        super().__init__()
        self.block_position_encoding = block_position_encoding
        self.word_embeddings = VocabEmbedding(vocab_size, hidden_size, padding_idx=padding_idx, init_method=init_method,
                                              amp_enabled=amp_enabled)
        n_pos = max_seq_length + 1 if block_position_encoding else max_seq_length
        self.position_embeddings = Embedding(n_pos, hidden_size, init_method=init_method, amp_enabled=amp_enabled)
        if block_position_encoding:
            self.block_position_embeddings = Embedding(n_pos, hidden_size, init_method=init_method, amp_enabled=amp_enabled)
        self.embedding_dropout = nn.Dropout(embedding_dropout_prob)

    def forward(self, input_ids, position_ids=None):
        b, s = input_ids.shape
        if position_ids is None:
            pos = torch.arange(s, device=input_ids.device)[None].expand(b, s)
            position_ids = torch.stack([pos, torch.zeros_like(pos)], 1) if self.block_position_encoding else pos
        x = self.word_embeddings(input_ids)
        if self.block_position_encoding:
            x = x + self.position_embeddings(position_ids[:, 0]) + self.block_position_embeddings(position_ids[:, 1])
        else:
            x = x + self.position_embeddings(position_ids)
        return self.embedding_dropout(x)


class GLMAttention(nn.Module):
    def __init__(self, hidden_size, num_attention_heads, attention_dropout_prob, output_dropout_prob, init_method,
                 output_layer_init_method, attention_scale, layer_idx):
        super().__init__()
        topo = dutil.get_dist_util()
        self.local_heads = num_attention_heads // topo.tensor_parallel_size
        self.head_size = hidden_size // num_attention_heads
        self.attention_scale = attention_scale
        self.query_key_value = Linear(hidden_size, 3 * hidden_size, parallel="col", init_method=init_method, layer_idx=layer_idx)
        self.dense = Linear(hidden_size, hidden_size, parallel="row", init_method=output_layer_init_method, layer_idx=layer_idx)
        self.attention_dropout_prob, self.output_dropout = attention_dropout_prob, nn.Dropout(output_dropout_prob)

    def forward(self, hidden, attention_mask, mem=None):
        b, q_len, _ = hidden.shape
        a, d = self.local_heads, self.head_size
        source = hidden if mem is None else torch.cat((mem, hidden), dim=1)
        qkv = self.query_key_value(source).view(b, -1, a, 3 * d).permute(0, 2, 1, 3)
        q, k, v = qkv[:, :, -q_len:, :d], qkv[..., d : 2 * d], qkv[..., 2 * d :]
        ctx = OF.attention(q, k, v, causal=False, scale=1.0 / math.sqrt(d), mask=attention_mask,
                           dropout_p=self.attention_dropout_prob, training=self.training)
        return self.output_dropout(self.dense(ctx.transpose(1, 2).reshape(b, q_len, a * d)))


class GLMLayer(nn.Module):
    def __init__(self, hidden_size, num_attention_heads, attention_dropout_prob, output_dropout_prob, layernorm_epsilon,
                 init_method, output_layer_init_method, attention_scale, layer_idx):
        super().__init__()
        self.layer_idx = layer_idx
        self.input_layernorm = LayerNorm(hidden_size, eps=layernorm_epsilon, layer_idx=layer_idx)
        self.attention = GLMAttention(hidden_size, num_attention_heads, attention_dropout_prob, output_dropout_prob,
                                      init_method, output_layer_init_method, attention_scale, layer_idx)
        self.post_attention_layernorm = LayerNorm(hidden_size, eps=layernorm_epsilon, layer_idx=layer_idx)
        self.mlp = MLP(hidden_size, 4 * hidden_size, output_dropout_prob, init_method, output_layer_init_method,
                       layer_idx=layer_idx)

    def forward(self, hidden, attention_mask, mem=None):
        ln = self.input_layernorm(hidden)
        mem = self.input_layernorm(mem) if mem is not None else None
        hidden = hidden + self.attention(ln, attention_mask, mem)
        return self.mlp(self.post_attention_layernorm(hidden), residual=hidden)


class Transformer(nn.Module):
    def __init__(self, num_layers, hidden_size, num_attention_heads, attention_dropout_prob=0.0, output_dropout_prob=0.0,
                 layernorm_epsilon=1e-5, init_method=None, output_layer_init_method=None, attention_scale=1.0):
        super().__init__()
        self.layers = nn.ModuleList([
            GLMLayer(hidden_size, num_attention_heads, attention_dropout_prob, output_dropout_prob, layernorm_epsilon,
                     init_method, output_layer_init_method, attention_scale, layer_idx=i) for i in range(num_layers)])
        self.final_layernorm = LayerNorm(hidden_size, eps=layernorm_epsilon, layer_idx=-1)

    def forward(self, hidden, attention_mask, memory_states=None):
        mem_layers = [hidden.detach()]
        for i, layer in enumerate(self.layers):
            hidden = layer(hidden, attention_mask, memory_states[i] if memory_states is not None else None)
            mem_layers.append(hidden.detach())
        return self.final_layernorm(hidden), mem_layers


class GLMModel(nn.Module):
    @configurable
    def __init__(self, num_layers, vocab_size, hidden_size, num_attention_heads, max_sequence_length=1024,
                 embedding_dropout_prob=0.0, attention_dropout_prob=0.0, output_dropout_prob=0.0, layernorm_epsilon=1e-5,
                 initializer_range=0.02, use_scaled_init_for_output_weights=True, bias_gelu_fusion=True,
                 bias_dropout_fusion=True, scale_mask_softmax_fusion=False, apply_query_key_layer_scaling=False,
                 amp_enabled=False, block_position_encoding=False, attention_scale=1.0, padding_idx=None, cfg=None):
        super().__init__()
        self.cfg = cfg
        init_method = init_method_normal(initializer_range)
        out_init = scaled_init_method_normal(initializer_range, num_layers) if use_scaled_init_for_output_weights else init_method
        self.embeddings = GLMEmbedding(vocab_size, hidden_size, max_sequence_length, padding_idx, init_method,
                                       embedding_dropout_prob, amp_enabled, block_position_encoding)
        self.transformer = Transformer(num_layers, hidden_size, num_attention_heads, attention_dropout_prob,
                                       output_dropout_prob, layernorm_epsilon, init_method, out_init, attention_scale)
        self.lm_head = LMLogits(vocab_size, bias=False)

    @classmethod
    def from_config(cls, cfg):
        keys = ("num_layers vocab_size hidden_size num_attention_heads max_sequence_length embedding_dropout_prob "
                "attention_dropout_prob output_dropout_prob layernorm_epsilon initializer_range "
                "use_scaled_init_for_output_weights bias_gelu_fusion bias_dropout_fusion scale_mask_softmax_fusion "
                "apply_query_key_layer_scaling amp_enabled block_position_encoding attention_scale padding_idx").split()
        out = {k: cfg[k] for k in keys if k in cfg}
        out["cfg"] = cfg
        return out

    @staticmethod
    def build_mask_matrix(batch_size, seq_length, sep, memory_length=0, device=None):
        """Prefix-LM visibility: everything before ``sep`` is bidirectional context, the rest is causal."""
        m = torch.ones(seq_length, seq_length, dtype=torch.bool, device=device).tril()[None].repeat(batch_size, 1, 1)
        sep_t = torch.as_tensor(sep, device=device).view(-1, 1).expand(batch_size, 1)
        ctx = torch.arange(seq_length, device=device)[None] < sep_t
        m = m | ctx[:, None, :]
        if memory_length > 0:
            m = torch.cat([torch.ones(batch_size, seq_length, memory_length, dtype=torch.bool, device=device), m], dim=2)
        return m[:, None]

    @staticmethod
    def update_mems(hiddens, mems):
        mem_len = mems[0].shape[1] if mems is not None else 0
        if mem_len == 0:
            return list(hiddens)
        return [torch.cat((m, h), dim=1) for m, h in zip(mems, hiddens)]

    def forward(self, input_ids, position_ids=None, attention_mask=None, memory_states=None, output_predict=True):
        b, q = input_ids.shape
        mem_len = memory_states[0].shape[1] if memory_states is not None else 0
        if attention_mask is None:
            attention_mask = torch.tensor(q, device=input_ids.device)
        if attention_mask.numel() == 1 or attention_mask.numel() == b and attention_mask.dim() <= 1:
            mask = self.build_mask_matrix(b, q, attention_mask, mem_len, input_ids.device)
        else:
            mask = attention_mask.bool()
            if mask.dim() == 2:
                mask = mask[:, None, None, :]
            mask = mask[..., -q - mem_len :]
        hidden, mem_layers = self.transformer(self.embeddings(input_ids, position_ids), mask, memory_states)
        mem_layers = self.update_mems(mem_layers, memory_states)
        if output_predict:
            hidden = self.lm_head(hidden, self.embeddings.word_embeddings.weight)
            if not self.training and dutil.get_dist_util().tensor_parallel_size > 1:
                hidden = mappings.gather_from_tp(hidden)
        return hidden, mem_layers


class GLMLoss(nn.Module):
    def forward(self, logits, labels):
        return {"lm_loss": torch.nn.functional.cross_entropy(logits.float().reshape(-1, logits.shape[-1]), labels.reshape(-1),
                                                             ignore_index=-100)}


class GLMForMultipleChoice(nn.Module):
    """Score every choice by the sum of the log-probs of its tokens (cloze-style classification)."""

    def __init__(self, cfg):
        super().__init__()
        self.glm = GLMModel(cfg)
        self.loss_func = GLMLoss()

    def forward(self, input_ids=None, position_ids=None, attention_mask=None, choice_ids=None, choice_indices=None,
                labels=None, mems=None, **kwargs):
        logits, _ = self.glm(input_ids, position_ids, attention_mask, mems)
        lp = torch.log_softmax(logits.float(), dim=-1)
        scores = []
        for out, choices, indices in zip(lp, choice_ids, choice_indices):
            scores.append(torch.stack([out[idx.to(out.device), ch.to(out.device)].sum() for ch, idx in zip(choices, indices)]))
        scores = torch.stack(scores)
        if labels is not None:
            return {"loss": torch.nn.functional.cross_entropy(scores, labels), "logits": scores}
        return {"logits": scores}


class GLMForConditionalGeneration(nn.Module, Generator):
    @configurable
    def __init__(self, cfg=None, **kwargs):
        super().__init__()
        self.cfg = cfg
        self.glm = GLMModel(cfg) if cfg is not None else GLMModel(**kwargs)
        self.loss_func = GLMLoss()
        self.past_key_values = [None]

    @classmethod
    def from_config(cls, cfg):
        return {"cfg": cfg}

    def forward(self, input_ids=None, position_ids=None, attention_mask=None, labels=None, memory_states=None, **kwargs):
        logits, mems = self.glm(input_ids, position_ids, attention_mask, memory_states)
        if labels is not None:
            return self.loss_func(logits, labels)
        return {"logits": logits, "past_key_values": mems}

    def set_cache(self, past):
        self.past_key_values = [None]

    def _reorder_cache(self, past, beam_idx):
        return None if past is None else [m.index_select(0, beam_idx.to(m.device)) for m in past]

    def _apply_reordered_cache(self, past):
        pass  # memories travel through ``model_kwargs["past"]``

    def prepare_inputs_for_generation(self, input_ids, past=None, position_ids=None, generation_attention_mask=None, **kwargs):
        mask, seq = generation_attention_mask, input_ids.shape[1]
        if past:
            if position_ids is not None:
                position_ids = position_ids[:, :, seq - 1].unsqueeze(-1)
            if mask is not None:
                mask = mask[:, :, seq - 1, :seq].unsqueeze(-2)
            input_ids = input_ids[:, -1:]
        else:
            if position_ids is not None:
                position_ids = position_ids[:, :, :seq]
            if mask is not None:
                mask = mask[:, :, :seq, :seq]
        return {"input_ids": input_ids, "position_ids": position_ids, "attention_mask": mask, "memory_states": past}

    @staticmethod
    def set_pipeline_stage_id(model):
        return model
