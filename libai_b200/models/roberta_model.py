"""RoBERTa (subclasses BERT).

Spec: reference libai/models/roberta_model.py — ``RobertaEmbeddings`` (:42-122; ``padding_idx`` on
word/position tables, position ids = cumulative count of non-pad tokens + ``pad_token_id``),
``RobertaLoss`` (:131-152), ``RobertaModel`` (:154-293), ``RobertaLMHead`` (:296-322; dense → GELU → LN
→ tied vocab-parallel logits with bias), ``RobertaForPreTraining`` (:396-449), ``RobertaForCausalLM``
(:452-512; next-token shifted loss).
"""
from __future__ import annotations

import torch
import torch.distributed as dist
from torch import nn

from libai_b200.config import configurable
from libai_b200.layers import Embedding, LayerNorm, Linear, LMLogits, ParallelCrossEntropyLoss, VocabEmbedding, build_activation
from libai_b200.layers._param import xavier_normal_
from libai_b200.layers.embedding import set_sp_shape
from libai_b200.parallel import mappings
from libai_b200.utils import distributed as dutil

from .bert_model import BertEmbeddings, BertExtendedAttnMask, BertModel, BertPooler
from .utils.pipeline_model import PipelineStageMixin
from .utils.weight_init import init_method_normal


class RobertaExtendedAttnMask(BertExtendedAttnMask):
    """Same as :class:`BertExtendedAttnMask`."""


class RobertaEmbeddings(BertEmbeddings):
    def __init__(self, vocab_size, hidden_size, max_sequence_length, embedding_dropout_prob, num_tokentypes=0,
                 pad_token_id=1, init_method=xavier_normal_, amp_enabled=False):
        super().__init__(vocab_size, hidden_size, max_sequence_length, embedding_dropout_prob,
                         num_tokentypes=num_tokentypes, init_method=init_method, amp_enabled=amp_enabled)
        self.pad_token_id = pad_token_id
        self.vocab_embeddings = VocabEmbedding(vocab_size, hidden_size, init_method=init_method,
                                               amp_enabled=amp_enabled, padding_idx=pad_token_id)
        self.position_embeddings = Embedding(max_sequence_length, hidden_size, init_method=init_method,
                                             amp_enabled=amp_enabled, padding_idx=pad_token_id)

    def forward(self, input_ids, tokentype_ids=None, position_ids=None):
        if position_ids is None:
            position_ids = self.create_position_ids_from_input_ids(input_ids, self.pad_token_id)
        return super().forward(input_ids, tokentype_ids, position_ids)

    @staticmethod
    def create_position_ids_from_input_ids(input_ids, pad_token_id):
        mask = input_ids.ne(pad_token_id).int()
        return (torch.cumsum(mask, dim=1).type_as(mask) * mask).long() + pad_token_id


class RobertaPooler(BertPooler):
    """Same as :class:`BertPooler`."""


class RobertaLoss(nn.Module):
    def __init__(self):
        super().__init__()
        self.lm_loss = ParallelCrossEntropyLoss()

    def forward(self, lm_output, lm_labels, loss_mask):
        per_token = self.lm_loss(lm_output, lm_labels)
        mask = loss_mask.float()
        denom = mask.sum()
        topo = dutil.get_dist_util()
        if topo.dp_group is not None:
            denom = denom.clone()
            dist.all_reduce(denom, group=topo.dp_group)
            denom = denom / topo.data_parallel_size
        return {"lm_loss": torch.sum(per_token.view(-1) * mask.view(-1)) / denom}


class RobertaModel(BertModel):
    @configurable
    def __init__(self, vocab_size, hidden_size, hidden_layers, num_attention_heads, intermediate_size,
                 hidden_dropout_prob, attention_probs_dropout_prob, max_position_embeddings, num_tokentypes=2,
                 add_pooling_layer=True, initializer_range=0.02, layernorm_eps=1e-12, pad_token_id=1,
                 bias_gelu_fusion=True, bias_dropout_fusion=True, scale_mask_softmax_fusion=True,
                 apply_query_key_layer_scaling=True, apply_residual_post_layernorm=False, amp_enabled=False):
        # BertModel.__init__ is @configurable-wrapped: call through with explicit keyword arguments
        super().__init__(
            vocab_size=vocab_size, hidden_size=hidden_size, hidden_layers=hidden_layers,
            num_attention_heads=num_attention_heads, intermediate_size=intermediate_size,
            hidden_dropout_prob=hidden_dropout_prob, attention_probs_dropout_prob=attention_probs_dropout_prob,
            max_position_embeddings=max_position_embeddings, num_tokentypes=num_tokentypes,
            add_pooling_layer=add_pooling_layer, initializer_range=initializer_range, layernorm_eps=layernorm_eps,
            bias_gelu_fusion=bias_gelu_fusion, bias_dropout_fusion=bias_dropout_fusion,
            scale_mask_softmax_fusion=scale_mask_softmax_fusion,
            apply_query_key_layer_scaling=apply_query_key_layer_scaling,
            apply_residual_post_layernorm=apply_residual_post_layernorm, amp_enabled=amp_enabled,
        )
        init_method = init_method_normal(initializer_range)
        self.embeddings = RobertaEmbeddings(vocab_size, hidden_size, max_position_embeddings, hidden_dropout_prob,
                                            num_tokentypes, pad_token_id, init_method, amp_enabled)
        self.extended_attn_mask = RobertaExtendedAttnMask()
        self.pooler = RobertaPooler(hidden_size, init_method) if add_pooling_layer else None

    @classmethod
    def from_config(cls, cfg):
        out = BertModel.from_config.__func__(cls, cfg)
        out["pad_token_id"] = cfg.pad_token_id
        return out

    def forward(self, input_ids, attention_mask, tokentype_ids=None, position_ids=None):
        batch = {"input_ids": input_ids, "tokentype_ids": tokentype_ids, "position_ids": position_ids,
                 "_ext_mask": self.extended_attn_mask(attention_mask)}
        return self.forward_stage(batch)

    def stage_pre(self, input_ids, tokentype_ids=None, position_ids=None, **_):
        return self.embeddings(input_ids, tokentype_ids, position_ids)


class RobertaLMHead(nn.Module):
    def __init__(self, vocab_size, hidden_size, init_method, layer_norm_eps):
        super().__init__()
        self.dense = Linear(hidden_size, hidden_size, bias=True, parallel="data", init_method=init_method, layer_idx=-1)
        self.activation_func = build_activation("gelu")
        self.layernorm = LayerNorm((hidden_size,), eps=layer_norm_eps, layer_idx=-1)
        self.lm_logits = LMLogits(vocab_size, bias=True)

    def forward(self, hidden_states, word_embeddings_weight):
        h = self.layernorm(self.dense(hidden_states, act="gelu"))
        topo = dutil.get_dist_util()
        if topo.tensor_parallel_size > 1:  # hidden_states are replicated over TP (gathered by the backbone)
            return torch.nn.functional.linear(mappings.copy_to_tp(h), word_embeddings_weight.to(h.dtype),
                                              self.lm_logits.bias.to(h.dtype))
        return self.lm_logits(h, word_embeddings_weight)


class RobertaPreTrainedModel(nn.Module, PipelineStageMixin):
    @staticmethod
    def set_pipeline_stage_id(model):
        return model

    @staticmethod
    def set_activation_checkpoint(model):
        model.activation_checkpoint = True
        return model

    # shared pipeline plumbing of the two heads below
    def forward_stage(self, batch, hidden_in=None):
        batch = dict(batch)
        batch["_ext_mask"] = self.roberta.extended_attn_mask(batch["attention_mask"])
        return PipelineStageMixin.forward_stage(self, batch, hidden_in)

    def stage_pre(self, input_ids, tokentype_ids=None, position_ids=None, **_):
        return self.roberta.embeddings(input_ids, tokentype_ids, position_ids)

    def stage_layers(self):
        return self.roberta.encoders

    def stage_layer_call(self, layer, hidden, batch):
        return layer(hidden, batch.get("_ext_mask"))


class RobertaForPreTraining(RobertaPreTrainedModel):
    def __init__(self, cfg):
        super().__init__()
        cfg.add_pooling_layer = False
        self.roberta = RobertaModel(cfg)
        self.lm_head = RobertaLMHead(cfg.vocab_size, cfg.hidden_size, init_method_normal(cfg.initializer_range), cfg.layernorm_eps)
        self.loss_fc = RobertaLoss()

    def forward(self, input_ids, attention_mask, tokentype_ids=None, lm_labels=None, loss_mask=None):
        return self.forward_stage(dict(input_ids=input_ids, attention_mask=attention_mask, tokentype_ids=tokentype_ids,
                                       lm_labels=lm_labels, loss_mask=loss_mask))

    def stage_post(self, hidden, lm_labels=None, loss_mask=None, **_):
        seq, _ = self.roberta.stage_post(hidden)
        scores = self.lm_head(seq, self.roberta.word_embeddings_weight())
        if lm_labels is not None:
            return self.loss_fc(scores, lm_labels, loss_mask)
        return {"prediction_scores": scores}


class RobertaForCausalLM(RobertaPreTrainedModel):
    def __init__(self, cfg):
        super().__init__()
        cfg.add_pooling_layer = False
        self.roberta = RobertaModel(cfg)
        self.lm_head = RobertaLMHead(cfg.vocab_size, cfg.hidden_size, init_method_normal(cfg.initializer_range), cfg.layernorm_eps)
        self.loss_fc = RobertaLoss()

    def forward(self, input_ids, attention_mask, tokentype_ids=None, position_ids=None, labels=None, loss_mask=None):
        return self.forward_stage(dict(input_ids=input_ids, attention_mask=attention_mask, tokentype_ids=tokentype_ids,
                                       position_ids=position_ids, labels=labels, loss_mask=loss_mask))

    def stage_post(self, hidden, labels=None, loss_mask=None, **_):
        seq, _ = self.roberta.stage_post(hidden)
        scores = self.lm_head(seq, self.roberta.word_embeddings_weight())
        if labels is not None:
            shifted = scores[:, :-1, :].contiguous()
            mask = loss_mask[:, 1:] if loss_mask is not None and loss_mask.shape[1] == labels.shape[1] else loss_mask
            if mask is None:
                mask = torch.ones_like(labels[:, 1:], dtype=torch.bool)
            return {"lm_loss": self.loss_fc(shifted, labels[:, 1:].contiguous(), mask)["lm_loss"]}
        return {"prediction_scores": scores}
