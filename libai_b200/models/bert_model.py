"""BERT.

Spec: reference libai/models/bert_model.py — ``BertExtendedAttnMask`` (:36-48), ``BertEmbeddings``
(:51-120), ``BertLMPredictionHead`` (:123-150; *replicated* dense → GELU → LN), ``BertPooler``
(:153-182; column-parallel on the [CLS] vector), ``BertLoss`` (:185-218; masked-LM loss = local
numerator / **global** mask count, SOP loss = plain mean), ``BertModel`` (:221-385),
``BertPreTrainingHeads`` (:388-423), ``BertForPreTraining`` (:426-554), ``BertForClassification``
(:557-591).  Parameter names match the reference.

Attention masking: the reference multiplies a ``[b,1,s,s]`` mask into the scores; here the 2-D padding
mask is turned into per-sample key lengths (``KeyPaddingMask``) that the flash-attention kernel
consumes directly (right-padded batches, which is what the BERT datasets produce); a dense mask is
still accepted and routed to the reference math.
"""
from __future__ import annotations

import torch
import torch.distributed as dist
from torch import nn

from libai_b200.layers.dropout import Dropout

from libai_b200.config import configurable
from libai_b200.layers import (
    Embedding,
    LayerNorm,
    Linear,
    LMLogits,
    ParallelCrossEntropyLoss,
    TransformerLayer,
    VocabEmbedding,
    build_activation,
)
from libai_b200.layers._param import create_parameter, xavier_normal_
from libai_b200.layers.attention import AttnMaskType, KeyPaddingMask
from libai_b200.layers.embedding import set_sp_shape
from libai_b200.parallel import mappings
from libai_b200.utils import distributed as dutil

from .utils.pipeline_model import PipelineStageMixin
from .utils.weight_init import init_method_normal, scaled_init_method_normal


class BertExtendedAttnMask(nn.Module):
    """``[b, s]`` padding mask → ``KeyPaddingMask`` (key lengths + lazily built ``[b,1,s,s]`` mask)."""

    def forward(self, attention_mask):
        return KeyPaddingMask(attention_mask)


class BertEmbeddings(nn.Module):
    def __init__(self, vocab_size, hidden_size, max_sequence_length, embedding_dropout_prob, num_tokentypes=0,
                 init_method=xavier_normal_, amp_enabled=False):
        super().__init__()
        self.vocab_embeddings = VocabEmbedding(vocab_size, hidden_size, init_method=init_method, amp_enabled=amp_enabled)
        self.position_embeddings = Embedding(max_sequence_length, hidden_size, init_method=init_method, amp_enabled=amp_enabled)
        self.tokentype_embeddings = (
            Embedding(num_tokentypes, hidden_size, init_method=init_method, amp_enabled=amp_enabled)
            if num_tokentypes > 0 else None
        )
        self.embedding_dropout = Dropout(embedding_dropout_prob)

    def forward(self, input_ids, tokentype_ids=None, position_ids=None):
        bsz, seq_length = input_ids.shape
        x = self.vocab_embeddings(input_ids)
        if position_ids is None:
            position_ids = torch.arange(seq_length, device=input_ids.device).unsqueeze(0)
        x = x + self.position_embeddings(position_ids)
        if self.tokentype_embeddings is not None:
            if tokentype_ids is None:
                tokentype_ids = torch.zeros_like(input_ids)
            x = x + self.tokentype_embeddings(tokentype_ids)
        x = self.embedding_dropout(x)
        if dutil.get_dist_util().sequence_parallel:
            set_sp_shape(bsz, seq_length)
            x = mappings.scatter_to_sp(x.reshape(-1, x.shape[-1]))
        return x

    def word_embeddings(self):
        return self.vocab_embeddings.weight


class BertLMPredictionHead(nn.Module):
    def __init__(self, hidden_size, init_method):
        super().__init__()
        self.dense = Linear(hidden_size, hidden_size, bias=True, parallel="data", init_method=init_method, layer_idx=-1)
        self.activation_func = build_activation("gelu")
        self.layernorm = LayerNorm((hidden_size,), layer_idx=-1)

    def forward(self, hidden_states):
        return self.layernorm(self.dense(hidden_states, act="gelu"))


class BertPooler(nn.Module):
    """tanh(W · h[CLS]) with a column-parallel ``W`` (output gathered so heads see the full vector)."""

    def __init__(self, hidden_size, init_method):
        super().__init__()
        self.dense = Linear(hidden_size, hidden_size, bias=True, parallel="col", init_method=init_method, layer_idx=-1)
        self.activation_func = build_activation("tanh")

    def forward(self, hidden_states):
        first = hidden_states[:, 0, :]
        topo = dutil.get_dist_util()
        if topo.tensor_parallel_size > 1:
            # the [CLS] vector is tiny: use the plain (non sequence-parallel) column-linear path
            y = torch.nn.functional.linear(mappings.copy_to_tp(first), self.dense.weight, self.dense.bias)
            y = mappings.gather_from_tp(y)
        else:
            y = self.dense(first)
        return self.activation_func(y)


class BertLoss(nn.Module):
    def __init__(self, add_binary_head):
        super().__init__()
        self.add_binary_head = add_binary_head
        self.lm_loss = ParallelCrossEntropyLoss()

    def forward(self, lm_output, lm_labels, loss_mask, binary_logits, ns_labels):
        per_token = self.lm_loss(lm_output, lm_labels)
        mask = loss_mask.float()
        # global-view semantics of the reference: the denominator counts masked tokens over the whole
        # (data-parallel) batch; with DP gradients averaged over D ranks this needs the D-scaled local share
        denom = mask.sum()
        topo = dutil.get_dist_util()
        if topo.dp_group is not None:
            denom = denom.clone()
            dist.all_reduce(denom, group=topo.dp_group)
            denom = denom / topo.data_parallel_size
        masked_lm_loss = torch.sum(per_token.view(-1) * mask.view(-1)) / (denom + 1e-7)
        out = {"lm_loss": masked_lm_loss}
        if self.add_binary_head:
            out["sop_loss"] = torch.nn.functional.cross_entropy(binary_logits.float(), ns_labels, ignore_index=-1)
        return out


class BertModel(nn.Module, PipelineStageMixin):
    """Bare BERT encoder: returns ``(sequence_output, pooled_output)``."""

    # `train.dist.sequence_parallel = "auto"` resolves to True for this model (token-sharded activations between
    # the tensor-parallel blocks are handled by its embeddings / heads)
    supports_sequence_parallel = True

    @configurable
    def __init__(self, vocab_size, hidden_size, hidden_layers, num_attention_heads, intermediate_size,
                 hidden_dropout_prob, attention_probs_dropout_prob, max_position_embeddings, num_tokentypes=2,
                 add_pooling_layer=True, initializer_range=0.02, layernorm_eps=1e-12, bias_gelu_fusion=True,
                 bias_dropout_fusion=True, scale_mask_softmax_fusion=True, apply_query_key_layer_scaling=True,
                 apply_residual_post_layernorm=False, amp_enabled=False):
        super().__init__()
        init_method = init_method_normal(initializer_range)
        scaled_init_method = scaled_init_method_normal(initializer_range, hidden_layers)
        self.hidden_size = hidden_size
        self.embeddings = BertEmbeddings(vocab_size, hidden_size, max_position_embeddings, hidden_dropout_prob,
                                         num_tokentypes, init_method, amp_enabled)
        self.extended_attn_mask = BertExtendedAttnMask()
        self.encoders = nn.ModuleList(
            [
                TransformerLayer(
                    hidden_size, intermediate_size, num_attention_heads,
                    attention_dropout_prob=attention_probs_dropout_prob, output_dropout_prob=hidden_dropout_prob,
                    layernorm_epsilon=layernorm_eps, bias_gelu_fusion=bias_gelu_fusion,
                    bias_dropout_fusion=bias_dropout_fusion, scale_mask_softmax_fusion=scale_mask_softmax_fusion,
                    apply_query_key_layer_scaling=apply_query_key_layer_scaling, init_method=init_method,
                    output_layer_init_method=scaled_init_method,
                    apply_residual_post_layernorm=apply_residual_post_layernorm,
                    attn_mask_type=AttnMaskType.padding, layer_idx=i,
                )
                for i in range(hidden_layers)
            ]
        )
        self.final_layernorm = LayerNorm((hidden_size,), eps=layernorm_eps, layer_idx=-1)
        self.pooler = BertPooler(hidden_size, init_method) if add_pooling_layer else None
        topo = dutil.get_dist_util()
        self.tied_weight_copy = None
        if topo.pipeline_parallel_size > 1:
            self.tied_weight_copy = create_parameter(
                (vocab_size, hidden_size), init_method, tp_dim=0, layer_idx=-1,
                shared_with=self.embeddings.vocab_embeddings.weight,
            )
            self.tied_weight_copy.shared_from = "embeddings.vocab_embeddings.weight"
            self.embeddings.vocab_embeddings.weight.is_tied_source = True

    @classmethod
    def from_config(cls, cfg):
        keys = (
            "vocab_size hidden_size hidden_layers num_attention_heads intermediate_size hidden_dropout_prob "
            "attention_probs_dropout_prob max_position_embeddings num_tokentypes add_pooling_layer initializer_range "
            "layernorm_eps bias_gelu_fusion bias_dropout_fusion scale_mask_softmax_fusion "
            "apply_query_key_layer_scaling apply_residual_post_layernorm amp_enabled"
        ).split()
        return {k: cfg[k] for k in keys}

    # ---- pipeline protocol ------------------------------------------------------------------------------
    def stage_pre(self, input_ids, tokentype_ids=None, **_):
        return self.embeddings(input_ids, tokentype_ids)

    def stage_layers(self):
        return self.encoders

    def stage_layer_call(self, layer, hidden, batch):
        return layer(hidden, batch.get("_ext_mask"))

    def stage_post(self, hidden, **_):
        topo = dutil.get_dist_util()
        seq = self.final_layernorm(hidden)
        if topo.sequence_parallel and seq.dim() == 2:
            from libai_b200.layers.embedding import get_sp_shape

            b, s = get_sp_shape()
            seq_full = mappings.gather_from_sp(seq, reduce_scatter_grad=True).view(b, s, -1)
        else:
            seq_full = seq
        pooled = self.pooler(seq_full) if self.pooler is not None else None
        return seq_full, pooled

    def forward(self, input_ids, attention_mask, tokentype_ids=None):
        batch = {"input_ids": input_ids, "tokentype_ids": tokentype_ids,
                 "_ext_mask": self.extended_attn_mask(attention_mask)}
        return self.forward_stage(batch)

    def word_embeddings_weight(self):
        topo = dutil.get_dist_util()
        if topo.pipeline_parallel_size > 1 and topo.is_last_stage and not topo.is_first_stage:
            return self.tied_weight_copy
        return self.embeddings.word_embeddings()


class BertPreTrainingHeads(nn.Module):
    def __init__(self, vocab_size, hidden_size, init_method, add_binary_head=True):
        super().__init__()
        self.predictions = BertLMPredictionHead(hidden_size, init_method)
        self.seq_relationship = Linear(hidden_size, 2, bias=True, parallel="data", init_method=init_method, layer_idx=-1)
        self.lm_logits = LMLogits(vocab_size, bias=True)
        self.loss_func = BertLoss(add_binary_head)

    def forward(self, sequence_output, pooled_output, word_embeddings_weight, ns_labels, lm_labels, loss_mask):
        prediction_scores = self.predictions(sequence_output)
        seq_relationship_score = self.seq_relationship(pooled_output) if pooled_output is not None else None
        topo = dutil.get_dist_util()
        # sequence_output is replicated over TP here (gathered in BertModel.stage_post)
        logits = torch.nn.functional.linear(
            mappings.copy_to_tp(prediction_scores), word_embeddings_weight.to(prediction_scores.dtype),
            None if self.lm_logits.bias is None else self.lm_logits.bias.to(prediction_scores.dtype),
        ) if topo.tensor_parallel_size > 1 else self.lm_logits(prediction_scores, word_embeddings_weight)
        if lm_labels is not None:
            return self.loss_func(logits, lm_labels, loss_mask, seq_relationship_score, ns_labels)
        return {"prediction_scores": logits, "seq_relationship_score": seq_relationship_score}


class BertForPreTraining(nn.Module, PipelineStageMixin):
    """BERT with the masked-LM head and the sentence-order (binary) head."""

    # `train.dist.sequence_parallel = "auto"` resolves to True for this model (token-sharded activations between
    # the tensor-parallel blocks are handled by its embeddings / heads)
    supports_sequence_parallel = True

    def __init__(self, cfg):
        super().__init__()
        self.bert = BertModel(cfg)
        self.cls_head = BertPreTrainingHeads(
            cfg.vocab_size, cfg.hidden_size, init_method_normal(cfg.initializer_range), cfg.add_binary_head
        )

    def forward(self, input_ids, attention_mask, tokentype_ids=None, ns_labels=None, lm_labels=None, loss_mask=None):
        batch = dict(input_ids=input_ids, attention_mask=attention_mask, tokentype_ids=tokentype_ids,
                     ns_labels=ns_labels, lm_labels=lm_labels, loss_mask=loss_mask)
        return self.forward_stage(batch)

    def forward_stage(self, batch, hidden_in=None):
        batch = dict(batch)
        batch["_ext_mask"] = self.bert.extended_attn_mask(batch["attention_mask"])
        return PipelineStageMixin.forward_stage(self, batch, hidden_in)

    def stage_pre(self, input_ids, tokentype_ids=None, **_):
        return self.bert.stage_pre(input_ids, tokentype_ids)

    def stage_layers(self):
        return self.bert.encoders

    def stage_layer_call(self, layer, hidden, batch):
        return layer(hidden, batch.get("_ext_mask"))

    def stage_post(self, hidden, ns_labels=None, lm_labels=None, loss_mask=None, **_):
        seq, pooled = self.bert.stage_post(hidden)
        return self.cls_head(seq, pooled, self.bert.word_embeddings_weight(), ns_labels, lm_labels, loss_mask)

    @staticmethod
    def set_pipeline_stage_id(model):
        """Placement is fixed at construction through ``layer_idx`` (API parity no-op)."""
        return model

    @staticmethod
    def set_activation_checkpoint(model):
        model.activation_checkpoint = True
        return model


class BertForClassification(nn.Module):
    """BERT + dropout + row-parallel classifier on the pooled output (reference :557-591)."""

    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        self.num_labels = cfg.num_labels
        self.bert = BertModel(cfg)
        self.classifier = Linear(
            cfg.hidden_size, cfg.num_labels, bias=True, parallel="data",
            init_method=init_method_normal(cfg.initializer_range), layer_idx=-1,
        )
        self.dropout = nn.Dropout(cfg.hidden_dropout_prob)
        self.loss_fct = nn.CrossEntropyLoss()

    def forward(self, input_ids, attention_mask, tokentype_ids=None, labels=None, **kwargs):
        _, pooled = self.bert(input_ids, attention_mask, tokentype_ids)
        logits = self.classifier(self.dropout(pooled)).view(-1, self.num_labels)
        if labels is not None:
            return {"loss": self.loss_fct(logits.float(), labels.view(-1))}
        return {"prediction_scores": logits}
