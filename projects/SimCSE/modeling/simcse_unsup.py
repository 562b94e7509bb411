"""Unsupervised SimCSE: the same sentence twice through dropout = positive pair; in-batch negatives; τ = 0.05
(reference projects/SimCSE/modeling/simcse_unsup.py)."""
import torch
from torch import nn

from libai_b200.config import configurable
from libai_b200.utils.checkpoint import Checkpointer

from .bert_for_simcse import BertForSimCSE
from .model_utils import MLPLayer, cosine_similarity


class _SimcseBase(nn.Module):
    group = 2

    def __init__(self, cfg):
        super().__init__()
        self.bert = BertForSimCSE(cfg)
        self.mlp = MLPLayer(cfg)
        self.pooler_type = cfg.pooler_type
        self.temp = cfg.get("temp", 0.05)
        weight = cfg.get("pretrained_model_weight", None)
        if weight is not None:
            from projects.SimCSE.utils.load_huggingface_weight import load_huggingface_bert

            load_huggingface_bert(self.bert, weight, cfg)

    def pooler(self, inputs, attention_mask):
        last, pooled, hidden = inputs
        mask = attention_mask.unsqueeze(-1).to(last.dtype)
        if self.pooler_type == "cls":
            return last[:, 0]
        if self.pooler_type == "pooled":
            return pooled
        if self.pooler_type == "last-avg":
            return (last * mask).sum(1) / mask.sum(1)
        if self.pooler_type == "first-last-avg":
            return ((hidden[1] + last) / 2.0 * mask).sum(1) / mask.sum(1)
        raise ValueError(self.pooler_type)

    def _encode(self, input_ids, attention_mask):
        bs, n = input_ids.shape[0], input_ids.shape[1]
        ids, mask = input_ids.reshape(bs * n, -1), attention_mask.reshape(bs * n, -1)
        return self.pooler(self.bert(ids, mask), mask), bs, n

    def _eval(self, input_ids, attention_mask, labels):
        out, bs, n = self._encode(input_ids, attention_mask)
        out = out.view(bs, n, -1)
        return {"sim": cosine_similarity(out[:, 0], out[:, 1]), "labels": labels}


class Simcse_unsup(_SimcseBase):
    def forward(self, input_ids, attention_mask, token_type_ids=None, labels=None):
        if not self.training:
            return self._eval(input_ids, attention_mask, labels)
        out, _, _ = self._encode(input_ids, attention_mask)
        out = self.mlp(out)
        idx = torch.arange(out.shape[0], device=out.device)
        target = (idx - idx % 2 * 2) + 1                      # 0↔1, 2↔3, …
        sim = cosine_similarity(out.unsqueeze(1), out.unsqueeze(0))
        sim = (sim - torch.eye(out.shape[0], device=out.device) * 1e12) / self.temp
        return {"loss": nn.functional.cross_entropy(sim, target)}
