"""BERT returning what the SimCSE poolers need: last hidden states, pooled output and the first layer's output
(reference projects/SimCSE/modeling/bert_for_simcse.py)."""
from libai_b200.models.bert_model import BertModel


class BertForSimCSE(BertModel):
    def forward(self, input_ids, attention_mask, tokentype_ids=None):
        first = []
        handle = self.encoders[0].register_forward_hook(lambda m, i, o: first.append(o[0] if isinstance(o, tuple) else o))
        try:
            seq, pooled = super().forward(input_ids, attention_mask, tokentype_ids)
        finally:
            handle.remove()
        return seq, pooled, (None, first[0] if first else None)
