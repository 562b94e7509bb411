"""Seq2Seq wrapper with the LM loss (reference projects/Couplets/modeling/model.py)."""
import torch
from torch import nn

from libai_b200.layers import ParallelCrossEntropyLoss
from projects.Couplets.modeling.transformer_model import TransformerModel


class Seq2SeqLoss(nn.Module):
    def __init__(self):
        super().__init__()
        self.lm_loss = ParallelCrossEntropyLoss()

    def forward(self, logits, lm_labels):
        per_token = self.lm_loss(logits, lm_labels)
        keep = (lm_labels != 0).float()
        return (per_token * keep).sum() / keep.sum().clamp(min=1.0)


class Seq2Seq(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.language_model = TransformerModel(cfg)
        self.loss_func = Seq2SeqLoss()

    def forward(self, encoder_input_ids, decoder_input_ids, encoder_attn_mask, decoder_attn_mask, encoder_decoder_attn_mask,
                lm_labels=None):
        logits = self.language_model(encoder_input_ids, decoder_input_ids, encoder_attn_mask, decoder_attn_mask,
                                     encoder_decoder_attn_mask)
        if lm_labels is not None:
            return {"total_loss": self.loss_func(logits, lm_labels)}
        return {"prediction_scores": logits}

    def encode(self, encoder_input_ids, encoder_attn_mask):
        return self.language_model.encode(encoder_input_ids, encoder_attn_mask)

    def decode(self, decoder_input_ids, decoder_attn_mask, encoder_states, encoder_decoder_attn_mask):
        return self.language_model.decode(decoder_input_ids, decoder_attn_mask, encoder_states, encoder_decoder_attn_mask)

    @staticmethod
    def set_pipeline_stage_id(model):
        return model
