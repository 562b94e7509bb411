"""LoRA adapters on the UNet's attention projections (reference projects/Stable_Diffusion/modeling.py:63-86 uses
diffusers' ``LoRACrossAttnProcessor`` + ``AttnProcsLayers``): every attention module gets rank-``r`` down/up pairs
for ``to_q``/``to_k``/``to_v``/``to_out``; only these train, and only these are saved."""
import torch
from torch import nn


class LoRAAttnAdapter(nn.Module):
    def __init__(self, hidden_size, cross_attention_dim=None, rank=4, scale=1.0):
        super().__init__()
        kv = cross_attention_dim or hidden_size
        self.rank, self.scale = rank, scale
        dims = {"to_q": (hidden_size, hidden_size), "to_k": (kv, hidden_size), "to_v": (kv, hidden_size),
                "to_out": (hidden_size, hidden_size)}
        for name, (din, dout) in dims.items():
            down, up = nn.Linear(din, rank, bias=False), nn.Linear(rank, dout, bias=False)
            nn.init.normal_(down.weight, std=1.0 / rank)
            nn.init.zeros_(up.weight)
            setattr(self, f"{name}_lora", nn.ModuleDict(dict(down=down, up=up)))

    def forward(self, name, x):
        pair = getattr(self, f"{name}_lora")
        w = pair["down"].weight
        return pair["up"](pair["down"](x.to(w.dtype))) * self.scale


class AttnProcsLayers(nn.Module):
    """All adapters of a UNet in one module: ``parameters()`` for the optimizer, ``state_dict()`` for saving."""

    def __init__(self, adapters: dict):
        super().__init__()
        self.keys_ = list(adapters)
        self.layers = nn.ModuleList([adapters[k] for k in self.keys_])

    def named_state(self):
        out = {}
        for key, layer in zip(self.keys_, self.layers):
            for n, p in layer.state_dict().items():
                out[f"{key}.processor.{n}"] = p
        return out


def add_lora_to_unet(unet, rank=4, scale=1.0) -> AttnProcsLayers:
    adapters = {}
    for name, attn in unet.attention_modules().items():
        hidden = attn.to_q.in_features
        cross = attn.to_k.in_features if name.endswith("attn2") else None
        adapter = LoRAAttnAdapter(hidden, cross, rank, scale).to(attn.to_q.weight.device)
        object.__setattr__(attn, "lora", adapter)     # not a submodule: the UNet's state_dict stays checkpoint-shaped
        adapters[name] = adapter
    return AttnProcsLayers(adapters)


def save_attn_procs(lora_layers: AttnProcsLayers, save_dir, name="pytorch_lora_weights.bin"):
    import os

    os.makedirs(save_dir, exist_ok=True)
    torch.save({k: v.detach().cpu() for k, v in lora_layers.named_state().items()}, os.path.join(save_dir, name))


def load_attn_procs(unet, path, rank=None, scale=1.0) -> AttnProcsLayers:
    import os

    file = os.path.join(path, "pytorch_lora_weights.bin") if os.path.isdir(path) else path
    state = torch.load(file, map_location="cpu", weights_only=True)
    if rank is None:
        rank = next(v.shape[0] for k, v in state.items() if k.endswith("down.weight"))
    layers = add_lora_to_unet(unet, rank, scale)
    for key, layer in zip(layers.keys_, layers.layers):
        prefix = f"{key}.processor."
        layer.load_state_dict({k[len(prefix):]: v for k, v in state.items() if k.startswith(prefix)})
    return layers
