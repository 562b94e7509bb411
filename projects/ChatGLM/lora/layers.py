"""LoRA layers (reference projects/ChatGLM/lora/layers.py: ``BaseTunerLayer`` / ``LoraLayer`` / ``Linear``).

``y = base(x) + (dropout(x) · Aᵀ · Bᵀ) · alpha/r`` with ``A ~ kaiming``, ``B = 0``; multiple named adapters, enable /
disable, ``merge`` (fold ``ΔW = B·A·scale`` into the base weight, optionally checked for NaNs) / ``unmerge``.
The base projection keeps running on the native GEMM; the rank-``r`` update is two skinny matmuls."""
from __future__ import annotations

import math
import warnings
from typing import Any, List, Optional

import torch
from torch import nn


def transpose(weight, fan_in_fan_out):
    return weight.T if fan_in_fan_out else weight


class LoraLayer:
    adapter_layer_names = ("lora_A", "lora_B")

    def _init_lora(self, base_layer: nn.Module):
        self.base_layer = base_layer
        self.r, self.lora_alpha, self.scaling = {}, {}, {}
        self.lora_dropout = nn.ModuleDict({})
        self.lora_A = nn.ParameterDict({})
        self.lora_B = nn.ParameterDict({})
        self._disable_adapters, self.merged_adapters = False, []
        self._active_adapter = "default"
        w = base_layer.weight
        self.out_features, self.in_features = w.shape[0], w.shape[1]

    def get_base_layer(self):
        base = self
        while hasattr(base, "base_layer"):
            base = base.base_layer
        return base

    @property
    def weight(self):
        return self.get_base_layer().weight

    @property
    def merged(self) -> bool:
        return bool(self.merged_adapters)

    @property
    def active_adapters(self) -> List[str]:
        a = self._active_adapter
        return [a] if isinstance(a, str) else list(a)

    def update_layer(self, adapter_name, r, lora_alpha, lora_dropout, init_lora_weights=True):
        if r <= 0:
            raise ValueError(f"`r` should be a positive integer value but the value passed is {r}")
        self.r[adapter_name], self.lora_alpha[adapter_name] = r, lora_alpha
        self.lora_dropout[adapter_name] = nn.Dropout(p=lora_dropout) if lora_dropout > 0.0 else nn.Identity()
        w = self.get_base_layer().weight
        self.lora_A[adapter_name] = nn.Parameter(torch.zeros(r, self.in_features, dtype=w.dtype, device=w.device))
        self.lora_B[adapter_name] = nn.Parameter(torch.zeros(self.out_features, r, dtype=w.dtype, device=w.device))
        self.scaling[adapter_name] = lora_alpha / r
        if init_lora_weights:
            self.reset_lora_parameters(adapter_name, init_lora_weights)

    def reset_lora_parameters(self, adapter_name, init_lora_weights=True):
        if init_lora_weights is True:
            nn.init.kaiming_uniform_(self.lora_A[adapter_name].float(), a=math.sqrt(5))
            with torch.no_grad():
                tmp = torch.empty_like(self.lora_A[adapter_name], dtype=torch.float32)
                nn.init.kaiming_uniform_(tmp, a=math.sqrt(5))
                self.lora_A[adapter_name].copy_(tmp)
        elif str(init_lora_weights).lower() == "gaussian":
            nn.init.normal_(self.lora_A[adapter_name], std=1 / self.r[adapter_name])
        nn.init.zeros_(self.lora_B[adapter_name])

    def set_scale(self, adapter, scale):
        if adapter in self.scaling:
            self.scaling[adapter] = scale * self.lora_alpha[adapter] / self.r[adapter]

    def scale_layer(self, scale: float):
        for a in self.active_adapters:
            if a in self.scaling and scale != 1:
                self.scaling[a] *= scale

    def unscale_layer(self, scale=None):
        for a in self.active_adapters:
            if a in self.scaling:
                self.scaling[a] = self.lora_alpha[a] / self.r[a] if scale is None else self.scaling[a] / scale

    def enable_adapters(self, enabled: bool):
        self._disable_adapters = not enabled
        for a in list(self.lora_A.keys()):
            self.lora_A[a].requires_grad_(enabled)
            self.lora_B[a].requires_grad_(enabled)

    def set_adapter(self, adapter_names):
        names = [adapter_names] if isinstance(adapter_names, str) else list(adapter_names)
        for a in self.lora_A.keys():
            self.lora_A[a].requires_grad_(a in names)
            self.lora_B[a].requires_grad_(a in names)
        self._active_adapter = names

    def delete_adapter(self, adapter_name: str):
        for store in (self.lora_A, self.lora_B, self.lora_dropout):
            if adapter_name in store:
                del store[adapter_name]
        for d in (self.r, self.lora_alpha, self.scaling):
            d.pop(adapter_name, None)
        if adapter_name in self.active_adapters:
            remaining = list(self.lora_A.keys())
            self._active_adapter = remaining[:1] if remaining else []
            if remaining:
                warnings.warn(f"Adapter {adapter_name} was active which is now deleted. Setting active adapter to {remaining[0]}.")


class Linear(nn.Module, LoraLayer):
    def __init__(self, base_layer, adapter_name: str, r: int = 0, lora_alpha: int = 1, lora_dropout: float = 0.0,
                 fan_in_fan_out: bool = False, init_lora_weights=True, **kwargs):
        super().__init__()
        self._init_lora(base_layer)
        self.fan_in_fan_out = fan_in_fan_out
        self._active_adapter = adapter_name
        self.update_layer(adapter_name, r, lora_alpha, lora_dropout, init_lora_weights)

    def get_delta_weight(self, adapter) -> torch.Tensor:
        a, b = self.lora_A[adapter], self.lora_B[adapter]
        return transpose((b.float() @ a.float()) * self.scaling[adapter], self.fan_in_fan_out).to(a.dtype)

    def merge(self, safe_merge: bool = False, adapter_names: Optional[List[str]] = None):
        if self.merged:
            warnings.warn(f"Already following adapters were merged {','.join(self.merged_adapters)}.")
        for a in (adapter_names or self.active_adapters):
            if a not in self.lora_A:
                continue
            base = self.get_base_layer()
            delta = self.get_delta_weight(a).to(base.weight.dtype)
            if safe_merge:
                merged = base.weight.data.clone() + delta
                if not torch.isfinite(merged).all():
                    raise ValueError(f"NaNs detected in the merged weights. The adapter {a} seems to be broken")
                base.weight.data.copy_(merged)
            else:
                base.weight.data.add_(delta)
            self.merged_adapters.append(a)

    def unmerge(self):
        if not self.merged:
            warnings.warn("Already unmerged. Nothing to do.")
            return
        while self.merged_adapters:
            a = self.merged_adapters.pop()
            if a in self.lora_A:
                self.get_base_layer().weight.data.sub_(self.get_delta_weight(a).to(self.weight.dtype))

    def forward(self, x, *args: Any, **kwargs: Any):
        if self._disable_adapters:
            if self.merged:
                self.unmerge()
            return self.base_layer(x, *args, **kwargs)
        result = self.base_layer(x, *args, **kwargs)
        if self.merged:
            return result
        extra = None
        for a in self.active_adapters:
            if a not in self.lora_A:
                continue
            h = self.lora_dropout[a](x).to(self.lora_A[a].dtype)
            upd = torch.nn.functional.linear(torch.nn.functional.linear(h, self.lora_A[a]), self.lora_B[a]) * self.scaling[a]
            extra = upd if extra is None else extra + upd
        if extra is None:
            return result
        if isinstance(result, tuple):  # skip_bias_add layers return (y, bias)
            return (result[0] + extra.to(result[0].dtype),) + tuple(result[1:])
        return result + extra.to(result.dtype)

    def __repr__(self) -> str:
        return "lora." + super().__repr__()


# The reference splits the adapter-layer protocol into ``BaseTunerLayer`` (generic) and ``LoraLayer`` (LoRA state); with
# LoRA as the only tuner here, one mixin carries both — the generic name is kept for isinstance checks in user code.
BaseTunerLayer = LoraLayer
