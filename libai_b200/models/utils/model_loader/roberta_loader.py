"""RoBERTa loaders (reference libai/models/utils/model_loader/roberta_loader.py): the BERT mapping with the
``roberta.`` prefix and the ``lm_head`` naming of HF RoBERTa."""
from .base_loader import ModelLoaderLiBai
from .bert_loader import BertLoaderHuggerFace


class RobertaLoaderHuggerFace(BertLoaderHuggerFace):
    hf_prefix = "roberta"

    def __init__(self, model, libai_cfg, pretrained_model_path, **kwargs):
        super().__init__(model, libai_cfg, pretrained_model_path, **kwargs)
        self.base_model_prefix_2 = "roberta"

    def _convert_state_dict(self, sd, cfg):
        out = super()._convert_state_dict(sd, cfg)
        return type(out)((("roberta." + k[5:]) if k.startswith("bert.") else k, v) for k, v in out.items())

    def _load_config_from_json(self, config_file):
        super()._load_config_from_json(config_file)
        cfg = self._read_config_json()
        self._map_config(cfg, {"pad_token_id": "pad_token_id"})


class RobertaLoaderLiBai(ModelLoaderLiBai):
    def __init__(self, model, libai_cfg, pretrained_model_path, **kwargs):
        super().__init__(model, libai_cfg, pretrained_model_path, **kwargs)
        self.base_model_prefix_2 = "roberta"
