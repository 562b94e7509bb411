"""Supervised SimCSE: (anchor, entailment, contradiction) triples; rows of anchors/positives are scored against
everything, hard negatives only appear as columns (reference projects/SimCSE/modeling/simcse_sup.py)."""
import torch
from torch import nn

from .model_utils import cosine_similarity
from .simcse_unsup import _SimcseBase


class Simcse_sup(_SimcseBase):
    group = 3

    @staticmethod
    def create_use_row(idx):
        return idx[idx % 3 != 2]

    def forward(self, input_ids, attention_mask, token_type_ids=None, labels=None):
        if not self.training:
            return self._eval(input_ids, attention_mask, labels)
        out, _, _ = self._encode(input_ids, attention_mask)
        out = self.mlp(out)
        idx = torch.arange(out.shape[0], device=out.device)
        use_row = self.create_use_row(idx)
        target = (use_row - use_row % 3 * 2) + 1              # anchor ↔ positive inside each triple
        sim = cosine_similarity(out.unsqueeze(1), out.unsqueeze(0))
        sim = (sim - torch.eye(out.shape[0], device=out.device) * 1e12)
        sim = sim.index_select(0, use_row) / self.temp
        return {"loss": nn.functional.cross_entropy(sim, target)}
