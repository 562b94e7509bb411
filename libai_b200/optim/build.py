"""Optimizer construction and per-parameter hyper-parameter groups.

Spec: reference libai/optim/build.py:32-162 — ``get_default_optimizer_params`` assigns weight-decay
overrides to normalisation layers and biases, supports per-name ``overrides``, stores
``clip_grad_max_norm`` / ``clip_grad_norm_type`` *inside* every param group, and merges groups
with identical hyper-parameters (few groups → few fused-optimizer launches).
"""
from __future__ import annotations

import copy
from collections import defaultdict
from typing import Any, Dict, List

import torch

from libai_b200.config import instantiate
from libai_b200.layers import LayerNorm, RMSLayerNorm

_NORM_TYPES = (
    LayerNorm,
    RMSLayerNorm,
    torch.nn.LayerNorm,
    torch.nn.BatchNorm1d,
    torch.nn.BatchNorm2d,
    torch.nn.BatchNorm3d,
    torch.nn.GroupNorm,
    torch.nn.InstanceNorm1d,
    torch.nn.InstanceNorm2d,
    torch.nn.InstanceNorm3d,
)


def build_optimizer(cfg, model):
    """``cfg`` is the lazy ``optim`` record; the model is injected into ``cfg.params.model``."""
    cfg.params.model = model
    return instantiate(cfg)


def get_default_optimizer_params(
    model,
    base_lr=None,
    weight_decay=None,
    weight_decay_norm=None,
    weight_decay_bias=None,
    clip_grad_max_norm=None,
    clip_grad_norm_type=None,
    overrides=None,
):
    overrides = dict(overrides or {})
    base: Dict[str, Any] = {}
    if base_lr is not None:
        base["lr"] = base_lr
    if weight_decay is not None:
        base["weight_decay"] = weight_decay
    if clip_grad_max_norm is not None and clip_grad_norm_type is not None:
        base["clip_grad_max_norm"] = clip_grad_max_norm
        base["clip_grad_norm_type"] = clip_grad_norm_type
    if weight_decay_bias is not None:
        if "bias" in overrides:
            raise ValueError("Conflicting overrides for 'bias'")
        overrides["bias"] = {"weight_decay": weight_decay_bias}

    per_param: List[Dict[str, Any]] = []
    seen = set()
    for module in model.modules():
        for pname, p in module.named_parameters(recurse=False):
            if not p.requires_grad or id(p) in seen:
                continue
            seen.add(id(p))
            if p.device.type == "meta":
                continue  # parameter of another pipeline stage
            hp = copy.copy(base)
            if isinstance(module, _NORM_TYPES) and weight_decay_norm is not None:
                hp["weight_decay"] = weight_decay_norm
            hp.update(overrides.get(pname, {}))
            per_param.append({"params": [p], **hp})
    return reduce_param_groups(per_param)


def _expand_param_groups(params: List[Dict[str, Any]]) -> List[Dict[str, Any]]:
    """One group per parameter; later items override hyper-parameters set by earlier ones."""
    table: Dict[Any, Dict[str, Any]] = {}
    order = []
    for item in params:
        assert "params" in item
        hp = {k: v for k, v in item.items() if k != "params"}
        for p in item["params"]:
            key = id(p) if isinstance(p, torch.Tensor) else p
            if key not in table:
                table[key] = {"params": [p]}
                order.append(key)
            table[key].update(hp)
    return [table[k] for k in order]


def reduce_param_groups(params: List[Dict[str, Any]]) -> List[Dict[str, Any]]:
    """Merge parameters whose hyper-parameters are identical into one group."""
    merged: Dict[tuple, list] = defaultdict(list)
    for item in _expand_param_groups(params):
        hp = tuple((k, v) for k, v in item.items() if k != "params")
        merged[hp].extend(item["params"])
    out = []
    for hp, plist in merged.items():
        g = {k: v for k, v in hp}
        g["params"] = plist
        out.append(g)
    return out


def set_weight_decay(model, skip_list=(), skip_keywords=()):
    """Two param groups – decayed / not decayed (1-D tensors, biases, names in ``skip_list`` or containing one of
    ``skip_keywords``) – as used by the SwinV2 recipes (reference configs/swinv2_imagenet.py:74-101 defines the
    same helper inline)."""
    if isinstance(skip_list, str):
        skip_list = (skip_list,)
    has_decay, no_decay = [], []
    for name, param in model.named_parameters():
        if not param.requires_grad or param.device.type == "meta":
            continue
        if param.dim() == 1 or name.endswith(".bias") or name in skip_list or any(k in name for k in skip_keywords):
            no_decay.append(param)
        else:
            has_decay.append(param)
    return [{"params": has_decay}, {"params": no_decay, "weight_decay": 0.0}]
