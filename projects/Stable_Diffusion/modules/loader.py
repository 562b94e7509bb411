"""Load / save the sub-models of a Stable Diffusion checkpoint in the public *diffusers* folder layout:

    <root>/unet/{config.json, diffusion_pytorch_model.safetensors|.bin}
    <root>/vae/{config.json, diffusion_pytorch_model.safetensors|.bin}
    <root>/scheduler/scheduler_config.json
    <root>/{tokenizer, text_encoder}/...            (transformers format)

Parameter names of ``modules/unet.py`` and ``modules/vae.py`` equal the checkpoint's; the only translations are
the pre-0.15 VAE attention names (``query/key/value/proj_attn``) and 1×1-conv vs linear ``proj_in/proj_out``
weights ([C_out, C_in, 1, 1] ↔ [C_out, C_in])."""
import json
import os

import torch

_OLD_VAE_ATTN = {"query": "to_q", "key": "to_k", "value": "to_v", "proj_attn": "to_out.0"}


def read_state(folder):
    st = os.path.join(folder, "diffusion_pytorch_model.safetensors")
    if os.path.exists(st):
        from safetensors.torch import load_file

        return load_file(st)
    return torch.load(os.path.join(folder, "diffusion_pytorch_model.bin"), map_location="cpu", weights_only=True)


def _adapt(state, model):
    want = model.state_dict()
    out = {}
    for k, v in state.items():
        parts = k.split(".")
        if len(parts) >= 2 and parts[-2] in _OLD_VAE_ATTN and "attentions" in parts:
            k = ".".join(parts[:-2] + [_OLD_VAE_ATTN[parts[-2]], parts[-1]])
        if k in want and want[k].shape != v.shape and want[k].numel() == v.numel():
            v = v.reshape(want[k].shape)
        out[k] = v
    return out


def load_submodel(cls, root, subfolder, strict=True, **overrides):
    folder = os.path.join(root, subfolder)
    with open(os.path.join(folder, "config.json")) as f:
        cfg = {k: v for k, v in json.load(f).items() if not k.startswith("_")}
    cfg.update(overrides)
    model = cls(**cfg)
    missing, unexpected = model.load_state_dict(_adapt(read_state(folder), model), strict=False)
    if strict and (missing or unexpected):
        raise RuntimeError(f"{subfolder}: missing keys {missing[:5]}… unexpected keys {unexpected[:5]}…")
    return model


def save_submodel(model, root, subfolder, class_name):
    from safetensors.torch import save_file

    folder = os.path.join(root, subfolder)
    os.makedirs(folder, exist_ok=True)
    with open(os.path.join(folder, "config.json"), "w") as f:
        json.dump({"_class_name": class_name, **model.config}, f, indent=2)
    save_file({k: v.detach().cpu().contiguous() for k, v in model.state_dict().items()},
              os.path.join(folder, "diffusion_pytorch_model.safetensors"))
