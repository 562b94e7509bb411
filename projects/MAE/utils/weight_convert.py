"""Official (PyTorch) MAE checkpoint → this project's parameter names (reference projects/MAE/utils/weight_convert.py):
``attn.qkv`` ([q; k; v]) → per-head interleaved ``self_attention.query_key_value``; ``mlp.fc1/fc2`` →
``dense_h_to_4h / dense_4h_to_h``; ``norm1/norm2`` → ``input_layernorm / post_attention_layernorm``."""
import re

import torch


def _fix_qkv(t, num_heads):
    head = t.shape[0] // 3 // num_heads
    if t.dim() == 2:
        return t.view(3, num_heads, head, t.shape[1]).permute(1, 0, 2, 3).reshape(-1, t.shape[1])
    return t.view(3, num_heads, head).permute(1, 0, 2).reshape(-1)


def convert_state_dict(torch_sd, num_heads, decoder_num_heads=None):
    out = {}
    rules = [(r"\.attn\.proj\.", ".self_attention.dense."), (r"\.norm1\.", ".input_layernorm."),
             (r"\.norm2\.", ".post_attention_layernorm."), (r"\.mlp\.fc1\.", ".mlp.dense_h_to_4h."),
             (r"\.mlp\.fc2\.", ".mlp.dense_4h_to_h.")]
    for k, v in torch_sd.items():
        if ".attn.qkv." in k:
            heads = decoder_num_heads if (k.startswith("decoder_blocks") and decoder_num_heads) else num_heads
            out[k.replace(".attn.qkv.", ".self_attention.query_key_value.")] = _fix_qkv(v, heads)
            continue
        for pat, rep in rules:
            k = re.sub(pat, rep, k)
        out[k] = v
    return out


def load_torch_checkpoint(model, path, strict=False, num_heads=12, decoder_num_heads=16):
    ckpt = torch.load(path, map_location="cpu", weights_only=False)
    sd = convert_state_dict(ckpt.get("model", ckpt), num_heads, decoder_num_heads)
    from libai_b200.parallel.state import load_full_state_dict

    return load_full_state_dict(model, sd, strict=strict)
