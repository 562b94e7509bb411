"""Load DALLE2-pytorch style checkpoints (LAION prior / decoder ``*.pth``) into this implementation.

Spec: reference projects/DALLE2/dalle2/dalle2_loader.py:21-96 — the reference broadcasts the torch state dict of
the prior and the decoder into its tensor-parallel modules.  Here each parameter is copied by (translated) name; a
tensor-parallel parameter (``tp_dim`` mark set by ``create_parameter``) receives this rank's slice.  This file knows
the systematic renames between the two code bases; anything it cannot place is *reported* (``strict=False``
returns the lists) instead of being silently dropped.
"""
import logging
import re

import torch

from libai_b200.utils import distributed as dist

logger = logging.getLogger(__name__)

# (regex on the checkpoint key, replacement) — DALLE2-pytorch 0.15 → this repo
_RENAMES = [
    (r"^net\.causal_transformer\.layers\.(\d+)\.0\.to_out\.0\.", r"net.causal_transformer.layers.\1.0.to_out."),
    (r"^net\.causal_transformer\.layers\.(\d+)\.0\.to_out\.1\.", r"net.causal_transformer.layers.\1.0.out_norm."),
    (r"^net\.causal_transformer\.layers\.(\d+)\.1\.0\.", r"net.causal_transformer.layers.\1.1.norm."),
    (r"^net\.causal_transformer\.layers\.(\d+)\.1\.3\.", r"net.causal_transformer.layers.\1.1.post_norm."),
    (r"^net\.causal_transformer\.layers\.(\d+)\.1\.5\.", r"net.causal_transformer.layers.\1.1.w_out."),
    (r"\.g$", ".weight"),
]


def _translate(key):
    for pat, rep in _RENAMES:
        key = re.sub(pat, rep, key)
    return key


def _shard(value, param):
    tp_dim = getattr(param, "tp_dim", None)
    topo = dist.get_dist_util()
    if tp_dim is None or topo.tensor_parallel_size == 1 or value.shape == param.shape:
        return value
    return value.chunk(topo.tensor_parallel_size, dim=tp_dim)[topo.tp_rank]


def load_state(module, state, strict=False):
    own = dict(module.named_parameters())
    own.update(dict(module.named_buffers()))
    loaded, unexpected = set(), []
    extra = {}
    for key, value in state.items():
        # the fused SwiGLU input projection of the reference layout = [value | gate] halves of ours
        m = re.match(r"^(net\.causal_transformer\.layers\.\d+\.1)\.1\.weight$", key)
        if m:
            v, g = value.chunk(2, dim=0)
            extra[f"{m.group(1)}.w_value.weight"], extra[f"{m.group(1)}.w_gate.weight"] = v, g
            continue
        extra[_translate(key)] = value
    with torch.no_grad():
        for key, value in extra.items():
            if key not in own:
                unexpected.append(key)
                continue
            value = _shard(value, own[key])
            if value.shape != own[key].shape:
                unexpected.append(f"{key} (shape {tuple(value.shape)} vs {tuple(own[key].shape)})")
                continue
            own[key].copy_(value.to(own[key].dtype))
            loaded.add(key)
    missing = [k for k in own if k not in loaded and not k.startswith("clip.")]
    if strict and (missing or unexpected):
        raise RuntimeError(f"missing: {missing[:8]} unexpected: {unexpected[:8]}")
    if missing or unexpected:
        logger.warning("checkpoint load: %d missing, %d unexpected keys", len(missing), len(unexpected))
    return missing, unexpected


class Dalle2ModelLoader:
    def __init__(self, model, libai_cfg=None, pretrained_model_path=None, **kwargs):
        self.model, self.cfg = model, libai_cfg

    def load(self):
        for part, path in (("prior", self.model.prior_weight_path), ("decoder", self.model.decoder_weight_path)):
            if not path:
                logger.warning("no %s checkpoint given: %s keeps its random initialisation", part, part)
                continue
            state = torch.load(path, map_location="cpu", weights_only=False)
            state = state.get("ema_model", state.get("model", state)) if isinstance(state, dict) else state
            load_state(getattr(self.model, part), state, strict=False)
        return self.model
