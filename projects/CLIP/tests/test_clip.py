"""CLIP numerics against ``transformers``' CLIPModel on random weights (the reference tests compare with the
OpenAI PyTorch implementation: projects/CLIP/tests/test_clip.py)."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..", "..")))


def _hf_to_openai(sd, layers_v, layers_t):
    out = {}
    out["visual.class_embedding"] = sd["vision_model.embeddings.class_embedding"]
    out["visual.conv1.weight"] = sd["vision_model.embeddings.patch_embedding.weight"]
    out["visual.positional_embedding"] = sd["vision_model.embeddings.position_embedding.weight"]
    out["visual.ln_pre.weight"], out["visual.ln_pre.bias"] = sd["vision_model.pre_layrnorm.weight"], sd["vision_model.pre_layrnorm.bias"]
    out["visual.ln_post.weight"], out["visual.ln_post.bias"] = sd["vision_model.post_layernorm.weight"], sd["vision_model.post_layernorm.bias"]
    out["visual.proj"] = sd["visual_projection.weight"].t()
    out["token_embedding.weight"] = sd["text_model.embeddings.token_embedding.weight"]
    out["positional_embedding"] = sd["text_model.embeddings.position_embedding.weight"]
    out["ln_final.weight"], out["ln_final.bias"] = sd["text_model.final_layer_norm.weight"], sd["text_model.final_layer_norm.bias"]
    out["text_projection"] = sd["text_projection.weight"].t()
    out["logit_scale"] = sd["logit_scale"]
    for src, dst, n in (("vision_model", "visual.transformer", layers_v), ("text_model", "transformer", layers_t)):
        for i in range(n):
            a, b = f"{src}.encoder.layers.{i}", f"{dst}.resblocks.{i}"
            out[f"{b}.attn.in_proj_weight"] = torch.cat([sd[f"{a}.self_attn.{p}_proj.weight"] for p in "qkv"])
            out[f"{b}.attn.in_proj_bias"] = torch.cat([sd[f"{a}.self_attn.{p}_proj.bias"] for p in "qkv"])
            for x, y in (("self_attn.out_proj", "attn.out_proj"), ("layer_norm1", "ln_1"), ("layer_norm2", "ln_2"),
                         ("mlp.fc1", "mlp.c_fc"), ("mlp.fc2", "mlp.c_proj")):
                out[f"{b}.{y}.weight"], out[f"{b}.{y}.bias"] = sd[f"{a}.{x}.weight"], sd[f"{a}.{x}.bias"]
    return out


def test_clip_matches_transformers():
    transformers = pytest.importorskip("transformers")
    from projects.CLIP.clip.model import build_model

    torch.manual_seed(0)
    cfg = transformers.CLIPConfig(
        text_config=dict(vocab_size=100, hidden_size=64, intermediate_size=256, num_hidden_layers=2, num_attention_heads=1,
                         max_position_embeddings=16, hidden_act="quick_gelu", eos_token_id=99, bos_token_id=98, pad_token_id=0),
        vision_config=dict(hidden_size=64, intermediate_size=256, num_hidden_layers=2, num_attention_heads=1, image_size=32,
                           patch_size=8, hidden_act="quick_gelu"),
        projection_dim=32)
    hf = transformers.CLIPModel(cfg).eval()
    model = build_model(_hf_to_openai(hf.state_dict(), 2, 2)).float()
    images = torch.randn(2, 3, 32, 32)
    text = torch.randint(1, 98, (2, 16))
    text[:, -1] = 99  # EOT = highest id → argmax picks it
    with torch.no_grad():
        ours, _ = model(images, text)
        theirs = hf(input_ids=text, pixel_values=images).logits_per_image
    assert (ours - theirs).abs().max() < 1e-3
