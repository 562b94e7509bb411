"""Vanilla encoder-decoder Transformer from the library layers (reference projects/Couplets/modeling/
transformer_model.py): vocab + sine position embeddings, N encoder ``TransformerLayer`` s, N decoder layers with cross
attention, tied LM head."""
import torch
from torch import nn

from libai_b200.config import configurable
from libai_b200.layers import LayerNorm, LMLogits, SinePositionalEmbedding, TransformerLayer, VocabEmbedding
from libai_b200.layers.attention import AttnMaskType
from libai_b200.models.utils.weight_init import init_method_normal, scaled_init_method_normal


class ExtendedMask(nn.Module):
    def forward(self, x):
        return x.unsqueeze(1)


class TransformerEmbedding(nn.Module):
    def __init__(self, vocab_size, hidden_size, max_sequence_length, embedding_dropout_prob, init_method=None):
        super().__init__()
        self.hidden_size = hidden_size
        self.word_embedding = VocabEmbedding(vocab_size, hidden_size, init_method=init_method)
        self.positional_encoding = SinePositionalEmbedding(max_sequence_length, hidden_size)
        self.embedding_dropout = nn.Dropout(embedding_dropout_prob)

    def forward(self, input_ids):
        pos = torch.arange(input_ids.shape[1], device=input_ids.device)
        x = self.word_embedding(input_ids) * (self.hidden_size ** 0.5)
        return self.embedding_dropout(x + self.positional_encoding(pos).to(x.dtype)[None])


def _stack(n, offset, is_decoder, cfg, init, out_init):
    return nn.ModuleList([
        TransformerLayer(cfg["hidden_size"], cfg["intermediate_size"], cfg["num_attention_heads"], is_decoder=is_decoder,
                         attention_dropout_prob=cfg["attention_dropout_prob"], output_dropout_prob=cfg["hidden_dropout_prob"],
                         layernorm_epsilon=cfg["layernorm_epsilon"], init_method=init, output_layer_init_method=out_init,
                         bias_gelu_fusion=cfg["bias_gelu_fusion"], bias_dropout_fusion=cfg["bias_dropout_fusion"],
                         scale_mask_softmax_fusion=cfg["scale_mask_softmax_fusion"],
                         apply_query_key_layer_scaling=cfg["apply_query_key_layer_scaling"],
                         attn_mask_type=AttnMaskType.padding, layer_idx=offset + i)
        for i in range(n)])


class TransformerModel(nn.Module):
    @configurable
    def __init__(self, vocab_size, max_position_embeddings, hidden_size=512, intermediate_size=512, hidden_layers=6,
                 num_attention_heads=8, embedding_dropout_prob=0.1, hidden_dropout_prob=0.1, attention_dropout_prob=0.1,
                 initializer_range=0.02, layernorm_epsilon=1e-5, bias_gelu_fusion=False, bias_dropout_fusion=False,
                 scale_mask_softmax_fusion=False, apply_query_key_layer_scaling=True):
        super().__init__()
        cfg = dict(hidden_size=hidden_size, intermediate_size=intermediate_size, num_attention_heads=num_attention_heads,
                   attention_dropout_prob=attention_dropout_prob, hidden_dropout_prob=hidden_dropout_prob,
                   layernorm_epsilon=layernorm_epsilon, bias_gelu_fusion=bias_gelu_fusion,
                   bias_dropout_fusion=bias_dropout_fusion, scale_mask_softmax_fusion=scale_mask_softmax_fusion,
                   apply_query_key_layer_scaling=apply_query_key_layer_scaling)
        init, out_init = init_method_normal(initializer_range), scaled_init_method_normal(initializer_range, hidden_layers)
        self.embedding = TransformerEmbedding(vocab_size, hidden_size, max_position_embeddings, embedding_dropout_prob, init)
        self.extended_attn_mask = ExtendedMask()
        self.encoder_layers = _stack(hidden_layers, 0, False, cfg, init, out_init)
        self.encoder_norm = LayerNorm(hidden_size, eps=layernorm_epsilon, layer_idx=hidden_layers - 1)
        self.decoder_layers = _stack(hidden_layers, hidden_layers, True, cfg, init, out_init)
        self.decoder_norm = LayerNorm(hidden_size, eps=layernorm_epsilon, layer_idx=-1)
        self.lm_head = LMLogits(vocab_size, bias=True)

    @classmethod
    def from_config(cls, cfg):
        return dict(cfg)

    def encode(self, encoder_input_ids, encoder_attn_mask):
        h, mask = self.embedding(encoder_input_ids), self.extended_attn_mask(encoder_attn_mask)
        for layer in self.encoder_layers:
            h = layer(h, mask)
        return self.encoder_norm(h)

    def decode(self, decoder_input_ids, decoder_attn_mask, encoder_states, encoder_decoder_attn_mask):
        h = self.embedding(decoder_input_ids)
        m1, m2 = self.extended_attn_mask(decoder_attn_mask), self.extended_attn_mask(encoder_decoder_attn_mask)
        for layer in self.decoder_layers:
            h = layer(h, m1, encoder_states, m2)
        return self.lm_head(self.decoder_norm(h), self.embedding.word_embedding.weight)

    def forward(self, encoder_input_ids, decoder_input_ids, encoder_attn_mask, decoder_attn_mask, encoder_decoder_attn_mask):
        enc = self.encode(encoder_input_ids, encoder_attn_mask)
        return self.decode(decoder_input_ids, decoder_attn_mask, enc, encoder_decoder_attn_mask)
