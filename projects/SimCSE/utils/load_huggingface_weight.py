"""Initialise the SimCSE BERT from an HF ``pytorch_model.bin`` (reference projects/SimCSE/utils/
load_huggingface_weight.py) — the conversion is the library's BERT loader."""
import os

import torch

from libai_b200.models.utils.model_loader.bert_loader import BertLoaderHuggerFace
from libai_b200.parallel.state import load_full_state_dict


def load_huggingface_bert(model, path, cfg):
    sd = torch.load(path, map_location="cpu", weights_only=True) if os.path.isfile(path) else None
    if sd is None:
        return BertLoaderHuggerFace(model, cfg, path).load()
    loader = BertLoaderHuggerFace(model, cfg, os.path.dirname(path))
    conv = loader._convert_state_dict(loader._fix_key(sd), cfg)
    conv = {k[5:] if k.startswith("bert.") else k: v for k, v in conv.items()}
    return load_full_state_dict(model, conv, strict=False)
