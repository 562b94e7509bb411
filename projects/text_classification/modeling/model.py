"""BERT + pooled-output classifier for GLUE / CLUE (reference projects/text_classification/modeling/model.py)."""
from torch import nn

from libai_b200.layers import Linear
from libai_b200.models.bert_model import BertModel
from libai_b200.models.utils.weight_init import init_method_normal

from .load_megatron_weight import load_megatron_bert


class ClassificationLoss(nn.Module):
    def forward(self, classification_logits, label):
        return nn.functional.cross_entropy(classification_logits.float(), label)


class ModelForSequenceClassification(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.num_classes = cfg.num_classes
        self.model = BertModel(cfg)
        if cfg.get("pretrain_megatron_weight") is not None:
            load_megatron_bert(self.model, cfg.pretrain_megatron_weight)
        self.loss_func = ClassificationLoss()
        self.classification_dropout = nn.Dropout(cfg.hidden_dropout_prob)
        self.classification_head = Linear(cfg.hidden_size, self.num_classes, bias=True, parallel="row",
                                          init_method=init_method_normal(cfg.initializer_range), layer_idx=-1)

    def forward(self, input_ids, attention_mask, token_type_ids=None, labels=None):
        _, pooled = self.model(input_ids, attention_mask, token_type_ids)
        logits = self.classification_head(self.classification_dropout(pooled))
        if self.training and labels is not None:
            return {"total_loss": self.loss_func(logits, labels)}
        return {"prediction_scores": logits}
