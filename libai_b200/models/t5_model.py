"""T5-style encoder-decoder (LiBai's Megatron variant: learned absolute positions, LayerNorm).

Spec: reference libai/models/t5_model.py — ``ExtendedMask`` (:33-35), ``T5Embedding`` (:38-83),
``T5Model`` (:86-345; encoder layers ``0…L-1``, decoder layers ``L…2L-1`` with cross attention,
encoder-state / KV cache for incremental decoding via ``set_cache``), ``T5Loss`` (:348-366; masked-LM
loss over the **global** mask count), ``T5ForPreTraining`` (:369-518).

Pipeline: the hidden state travelling between stages is the pair ``(encoder_states, decoder_states)``
(the decoder stream is empty until the decoder stages), mirroring the reference's
``pipeline_num_layers = 2 · hidden_layers`` layout (configs/t5_large_pretrain.py:26).
"""
from __future__ import annotations

import torch
import torch.distributed as dist
from torch import nn

from libai_b200.layers.dropout import Dropout

from libai_b200.config import configurable
from libai_b200.layers import Embedding, LayerNorm, LMLogits, ParallelCrossEntropyLoss, TransformerLayer, VocabEmbedding
from libai_b200.layers._param import create_parameter, xavier_normal_
from libai_b200.layers.attention import AttnMaskType
from libai_b200.utils import distributed as dutil

from .utils.pipeline_model import PipelineStageMixin
from .utils.weight_init import init_method_normal, scaled_init_method_normal


class ExtendedMask(nn.Module):
    def forward(self, attention_mask):
        return attention_mask.unsqueeze(1)


class T5Embedding(nn.Module):
    def __init__(self, hidden_size, vocab_size, max_sequence_length, embedding_dropout_prob, init_method=xavier_normal_,
                 amp_enabled=False):
        super().__init__()
        self.hidden_size, self.vocab_size = hidden_size, vocab_size
        self.word_embeddings = VocabEmbedding(num_embeddings=vocab_size, embedding_dim=hidden_size,
                                              init_method=init_method, amp_enabled=amp_enabled)
        self.position_embeddings = Embedding(num_embeddings=max_sequence_length, embedding_dim=hidden_size,
                                             init_method=init_method, amp_enabled=amp_enabled)
        self.embedding_dropout = Dropout(embedding_dropout_prob)

    def forward(self, input_ids, past_length=0):
        seq_length = input_ids.shape[1]
        position_ids = torch.arange(past_length, past_length + seq_length, device=input_ids.device).unsqueeze(0)
        x = self.word_embeddings(input_ids) + self.position_embeddings(position_ids)
        return self.embedding_dropout(x)


class T5Model(nn.Module, PipelineStageMixin):
    @configurable
    def __init__(self, vocab_size, hidden_size, hidden_layers, num_attention_heads, intermediate_size,
                 embedding_dropout_prob, hidden_dropout_prob, attention_probs_dropout_prob, max_position_embeddings,
                 initializer_range=0.02, layernorm_eps=1e-12, bias_gelu_fusion=False, bias_dropout_fusion=False,
                 scale_mask_softmax_fusion=False, apply_query_key_layer_scaling=True,
                 apply_residual_post_layernorm=False, amp_enabled=False):
        super().__init__()
        init_method = init_method_normal(initializer_range)
        scaled_init_method = scaled_init_method_normal(initializer_range, hidden_layers)
        self.hidden_layers = hidden_layers
        self.embedding = T5Embedding(hidden_size=hidden_size, vocab_size=vocab_size,
                                     max_sequence_length=max_position_embeddings,
                                     embedding_dropout_prob=embedding_dropout_prob, init_method=init_method,
                                     amp_enabled=amp_enabled)
        self.extended_attn_mask = ExtendedMask()

        def block(i, is_decoder):
            return TransformerLayer(
                hidden_size=hidden_size, ffn_hidden_size=intermediate_size, num_attention_heads=num_attention_heads,
                is_decoder=is_decoder, attention_dropout_prob=attention_probs_dropout_prob,
                output_dropout_prob=hidden_dropout_prob, layernorm_epsilon=layernorm_eps, init_method=init_method,
                output_layer_init_method=scaled_init_method, bias_gelu_fusion=bias_gelu_fusion,
                bias_dropout_fusion=bias_dropout_fusion, scale_mask_softmax_fusion=scale_mask_softmax_fusion,
                apply_query_key_layer_scaling=apply_query_key_layer_scaling,
                apply_residual_post_layernorm=apply_residual_post_layernorm,
                attn_mask_type=AttnMaskType.padding, layer_idx=i,
            )

        self.encoder = nn.Sequential()
        self.encoder.add_module("layers", nn.ModuleList([block(i, False) for i in range(hidden_layers)]))
        self.encoder.add_module("final_layernorm", LayerNorm((hidden_size,), eps=layernorm_eps, layer_idx=hidden_layers - 1))
        self.decoder = nn.Sequential()
        self.decoder.add_module("layers", nn.ModuleList([block(i, True) for i in range(hidden_layers, 2 * hidden_layers)]))
        self.decoder.add_module("final_layernorm", LayerNorm((hidden_size,), eps=layernorm_eps, layer_idx=2 * hidden_layers - 1))
        self.past_key_values = [None] * len(self.decoder.layers)
        self.encoder_states = None
        self.past_length = 0
        self.lm_head = LMLogits(vocab_size, bias=True)
        # tied embedding under pipeline parallelism: see GPTModel (the last stage keeps a synchronised copy)
        topo = dutil.get_dist_util()
        self.tied_weight_copy = None
        if topo.pipeline_parallel_size > 1:
            self.tied_weight_copy = create_parameter(
                (vocab_size, hidden_size), init_method, tp_dim=0, layer_idx=-1,
                shared_with=self.embedding.word_embeddings.weight,
            )
            self.tied_weight_copy.shared_from = "embedding.word_embeddings.weight"
            self.embedding.word_embeddings.weight.is_tied_source = True

    def word_embeddings_weight(self):
        topo = dutil.get_dist_util()
        if topo.pipeline_parallel_size > 1 and topo.is_last_stage and not topo.is_first_stage:
            return self.tied_weight_copy
        return self.embedding.word_embeddings.weight

    # ---- pipeline protocol -------------------------------------------------------------------------
    # The running "hidden state" of an encoder-decoder model is the pair (encoder stream, decoder stream).  Both
    # embeddings live on the first stage, so the decoder-input embedding rides along through the encoder stages;
    # from the first decoder layer on, the first slot holds the (final-normed) encoder output that every decoder
    # layer cross-attends to.  Layer indices 0..L-1 are encoder blocks, L..2L-1 decoder blocks
    # (train.dist.pipeline_num_layers = 2 * hidden_layers).
    def stage_pre(self, encoder_input_ids, decoder_input_ids, **_):
        return self.embedding(encoder_input_ids), self.embedding(decoder_input_ids, 0)

    def stage_layers(self):
        return list(self.encoder.layers) + list(self.decoder.layers)

    def stage_layer_call(self, layer, hidden, batch):
        enc, dec = hidden
        if layer.layer_idx < self.hidden_layers:
            enc = layer(enc, self.extended_attn_mask(batch["encoder_attn_mask"]))
            if layer.layer_idx == self.hidden_layers - 1:
                enc = self.encoder.final_layernorm(enc)
        else:
            dec = layer(dec, self.extended_attn_mask(batch["decoder_attn_mask"]), enc,
                        self.extended_attn_mask(batch["encoder_decoder_attn_mask"]))
        return enc, dec

    def stage_post(self, hidden, **_):
        _, dec = hidden
        return self.lm_head(self.decoder.final_layernorm(dec), self.word_embeddings_weight())

    @classmethod
    def from_config(cls, cfg):
        keys = (
            "vocab_size hidden_size hidden_layers num_attention_heads intermediate_size embedding_dropout_prob "
            "hidden_dropout_prob attention_probs_dropout_prob max_position_embeddings initializer_range layernorm_eps "
            "bias_gelu_fusion bias_dropout_fusion scale_mask_softmax_fusion apply_query_key_layer_scaling "
            "apply_residual_post_layernorm amp_enabled"
        ).split()
        return {k: cfg[k] for k in keys}

    def forward(self, encoder_input_ids, decoder_input_ids, encoder_attn_mask, decoder_attn_mask,
                encoder_decoder_attn_mask, use_cache=False):
        """Returns the (vocab-split) logits ``[b, s_dec, V/t]``."""
        if use_cache and self.encoder_states is not None:
            encoder_states = self.encoder_states
        else:
            self.set_cache(encoder_states=None, past_key_values=None)
            enc_mask = self.extended_attn_mask(encoder_attn_mask)
            h = self.embedding(encoder_input_ids)
            for layer in self.encoder.layers:
                h = layer(h, enc_mask)
            encoder_states = self.encoder.final_layernorm(h)
        dec_mask = self.extended_attn_mask(decoder_attn_mask)
        cross_mask = self.extended_attn_mask(encoder_decoder_attn_mask)
        h = self.embedding(decoder_input_ids, self.past_length)
        presents = []
        for layer, past in zip(self.decoder.layers, self.past_key_values):
            h = layer(h, dec_mask, encoder_states, cross_mask, past_key_value=past, use_cache=use_cache)
            if use_cache:
                h, present = h
                presents.append(present)
        if use_cache:
            self.set_cache(encoder_states, past_key_values=presents)
        decoder_states = self.decoder.final_layernorm(h)
        return self.lm_head(decoder_states, self.word_embeddings_weight())

    def set_cache(self, encoder_states, past_key_values):
        self.encoder_states = encoder_states
        self.past_length = 0 if past_key_values is None else past_key_values[0][0].shape[2]
        if past_key_values is None:
            past_key_values = [None] * len(self.decoder.layers)
        assert len(past_key_values) == len(self.decoder.layers), (
            f"past_key_values's length {len(past_key_values)} doesn't match "
            f"decoder num_layers' length {len(self.decoder.layers)}"
        )
        self.past_key_values = past_key_values


class T5Loss(nn.Module):
    def __init__(self) -> None:
        super().__init__()
        self.lm_loss = ParallelCrossEntropyLoss()

    def forward(self, logits, lm_labels, loss_mask):
        per_token = self.lm_loss(logits, lm_labels)
        mask = loss_mask.float()
        denom = mask.sum()
        topo = dutil.get_dist_util()
        if topo.dp_group is not None:  # global mask count (see BertLoss)
            denom = denom.clone()
            dist.all_reduce(denom, group=topo.dp_group)
            denom = denom / topo.data_parallel_size
        return {"masked_lm_loss": torch.sum(per_token.view(-1) * mask.view(-1)) / denom}


class T5ForPreTraining(nn.Module, PipelineStageMixin):
    """T5 with the span-corruption LM loss."""

    def __init__(self, cfg) -> None:
        super().__init__()
        self.t5_model = T5Model(cfg)
        self.loss_func = T5Loss()

    def set_cache(self, encoder_states, past_key_values):
        self.t5_model.set_cache(encoder_states, past_key_values)

    def forward(self, encoder_input_ids, decoder_input_ids, encoder_attn_mask, decoder_attn_mask,
                encoder_decoder_attn_mask, lm_labels=None, loss_mask=None, use_cache=False):
        logits = self.t5_model(encoder_input_ids, decoder_input_ids, encoder_attn_mask, decoder_attn_mask,
                               encoder_decoder_attn_mask, use_cache=use_cache)
        if lm_labels is not None:
            return self.loss_func(logits, lm_labels, loss_mask)
        return {"prediction_scores": logits}

    # pipeline protocol: delegate to the backbone, add the loss on the last stage
    def stage_pre(self, **batch):
        return self.t5_model.stage_pre(**batch)

    def stage_layers(self):
        return self.t5_model.stage_layers()

    def stage_layer_call(self, layer, hidden, batch):
        return self.t5_model.stage_layer_call(layer, hidden, batch)

    def stage_post(self, hidden, lm_labels=None, loss_mask=None, **_):
        logits = self.t5_model.stage_post(hidden)
        if lm_labels is not None:
            return self.loss_func(logits, lm_labels, loss_mask)
        return {"prediction_scores": logits}

    @staticmethod
    def set_pipeline_stage_id(model):
        """API parity: placement is decided at construction through each layer's ``layer_idx``."""
        return model

    @staticmethod
    def set_activation_checkpoint(model):
        model.activation_checkpoint = True
        return model
