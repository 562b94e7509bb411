"""Initialise ``BertModel`` from a Megatron-LM BERT checkpoint (reference projects/text_classification/modeling/
load_megatron_weight.py): ``model.language_model.{embedding,transformer,pooler}`` → library names; Megatron's fused
qkv already uses the per-head interleaved layout."""
import logging

import torch

from libai_b200.parallel.state import load_full_state_dict

logger = logging.getLogger(__name__)


def convert_megatron_state(lm):
    out = {}
    emb = lm["embedding"]
    out["embeddings.vocab_embeddings.weight"] = emb["word_embeddings"]["weight"]
    out["embeddings.position_embeddings.weight"] = emb["position_embeddings"]["weight"]
    if "tokentype_embeddings" in emb:
        out["embeddings.tokentype_embeddings.weight"] = emb["tokentype_embeddings"]["weight"]
    rename = {"attention.query_key_value": "self_attention.query_key_value", "attention.dense": "self_attention.dense"}
    enc = lm.get("transformer", lm.get("encoder"))
    for k, v in enc.items():
        if k.startswith("final_layernorm"):
            out[k] = v
            continue
        k2 = k.replace("layers.", "encoders.", 1)
        for a, b in rename.items():
            k2 = k2.replace(a, b)
        out[k2] = v
    for k, v in lm.get("pooler", {}).items():
        out["pooler." + k] = v
    return out


def load_megatron_bert(model, path):
    ckpt = torch.load(path, map_location="cpu", weights_only=False)
    lm = ckpt["model"]["language_model"] if "model" in ckpt else ckpt["language_model"]
    missing, unexpected, mismatched = load_full_state_dict(model, convert_megatron_state(lm), strict=False)
    logger.info(f"megatron weights loaded: missing={missing} unexpected={unexpected} mismatched={mismatched}")
    return model
