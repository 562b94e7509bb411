"""``LoraModel``: inject LoRA adapters into a model by module-name matching.

Spec: reference projects/ChatGLM/lora/lora_model.py — ``BaseTuner.inject_adapter`` (:131-193), trainability
marking with the ``bias`` policy ``none | all | lora_only`` (:342-361), per-module rank/alpha patterns (:290-300),
``merge_and_unload`` / ``unload`` (:423-497), adapter enable / disable / set / delete, and a ``state_dict`` that
only carries the adapter tensors (:500-510)."""
from __future__ import annotations

import re
from typing import Any, List, Optional

import torch
from torch import nn

from .layers import Linear as LoraLinear
from .layers import LoraLayer
from .utils import _get_submodules, check_target_module_exists


def _cfg_get(cfg, key, default=None):
    if hasattr(cfg, "get"):
        value = cfg.get(key, default)
        return default if value is None else value
    return getattr(cfg, key, default)


class LoraModel(nn.Module):
    prefix = "lora_"

    def __init__(self, model, config, adapter_name: str = "default") -> None:
        super().__init__()
        self.model = model
        self.peft_config = {adapter_name: config}
        self.active_adapter = adapter_name
        self.inject_adapter(self.model, adapter_name)

    def forward(self, *args: Any, **kwargs: Any):
        return self.model(*args, **kwargs)

    def __getattr__(self, name: str):
        try:
            return super().__getattr__(name)
        except AttributeError:
            return getattr(self.model, name)

    # ------------------------------------------------------------------ injection
    def _create_and_replace(self, cfg, adapter_name, target, target_name, parent, current_key):
        rank_pattern = dict(_cfg_get(cfg, "rank_pattern", {}) or {})
        alpha_pattern = dict(_cfg_get(cfg, "alpha_pattern", {}) or {})
        key = next((k for k in list(rank_pattern) + list(alpha_pattern) if re.match(rf".*\.{k}$", current_key)), current_key)
        r = rank_pattern.get(key, _cfg_get(cfg, "r", 8))
        alpha = alpha_pattern.get(key, _cfg_get(cfg, "lora_alpha", 8))
        if isinstance(target, LoraLayer):
            target.update_layer(adapter_name, r, alpha, _cfg_get(cfg, "lora_dropout", 0.0), _cfg_get(cfg, "init_lora_weights", True))
            return
        if not hasattr(target, "weight") or target.weight.dim() != 2:
            raise ValueError(f"Target module {target} is not supported. Only linear layers are supported.")
        new = LoraLinear(target, adapter_name, r=r, lora_alpha=alpha, lora_dropout=_cfg_get(cfg, "lora_dropout", 0.0),
                         fan_in_fan_out=_cfg_get(cfg, "fan_in_fan_out", False),
                         init_lora_weights=_cfg_get(cfg, "init_lora_weights", True))
        setattr(parent, target_name, new)

    def inject_adapter(self, model: nn.Module, adapter_name: str):
        cfg = self.peft_config[adapter_name]
        found = False
        for key in [k for k, _ in model.named_modules()]:
            if not key or not check_target_module_exists(cfg, key):
                continue
            parent, target, target_name = _get_submodules(model, key)
            if isinstance(target, LoraLinear) and adapter_name in target.lora_A:
                continue
            found = True
            self._create_and_replace(cfg, adapter_name, target, target_name, parent, current_key=key)
        if not found:
            raise ValueError(f"Target modules {_cfg_get(cfg, 'target_modules')} not found in the base model. "
                             "Please check the target modules and try again.")
        self._mark_only_adapters_as_trainable(model)
        if _cfg_get(cfg, "inference_mode", False):
            for n, p in model.named_parameters():
                if adapter_name in n:
                    p.requires_grad = False

    def _mark_only_adapters_as_trainable(self, model: nn.Module) -> None:
        for n, p in model.named_parameters():
            if self.prefix not in n:
                p.requires_grad = False
        for adapter in self.peft_config:
            bias = _cfg_get(self.peft_config[adapter], "bias", "none")
            if bias == "none":
                continue
            if bias == "all":
                for n, p in model.named_parameters():
                    if "bias" in n:
                        p.requires_grad = True
            elif bias == "lora_only":
                for m in model.modules():
                    if isinstance(m, LoraLayer) and getattr(m.get_base_layer(), "bias", None) is not None:
                        m.get_base_layer().bias.requires_grad = True
            else:
                raise NotImplementedError(f"Requested bias: {bias}, is not implemented.")

    # ------------------------------------------------------------------ adapter management
    def _lora_layers(self):
        return [m for m in self.model.modules() if isinstance(m, LoraLayer)]

    def enable_adapter_layers(self) -> None:
        for m in self._lora_layers():
            m.enable_adapters(True)

    def disable_adapter_layers(self) -> None:
        for m in self._lora_layers():
            m.enable_adapters(False)

    def set_adapter(self, adapter_name) -> None:
        for m in self._lora_layers():
            if m.merged:
                m.unmerge()
            m.set_adapter(adapter_name)
        self.active_adapter = adapter_name

    def delete_adapter(self, adapter_name: str) -> None:
        if adapter_name not in self.peft_config:
            raise ValueError(f"Adapter {adapter_name} does not exist")
        del self.peft_config[adapter_name]
        for m in self._lora_layers():
            m.delete_adapter(adapter_name)

    def merge_adapter(self, safe_merge=False, adapter_names: Optional[List[str]] = None) -> None:
        for m in self._lora_layers():
            m.merge(safe_merge=safe_merge, adapter_names=adapter_names)

    def unmerge_adapter(self):
        for m in self._lora_layers():
            m.unmerge()

    def _unload_and_optionally_merge(self, merge=True, safe_merge=False, adapter_names=None):
        for key in [k for k, _ in self.model.named_modules() if self.prefix not in k]:
            try:
                parent, target, target_name = _get_submodules(self.model, key)
            except AttributeError:
                continue
            if isinstance(target, LoraLinear):
                if merge:
                    target.merge(safe_merge=safe_merge, adapter_names=adapter_names)
                setattr(parent, target_name, target.get_base_layer())
        return self.model

    def merge_and_unload(self, safe_merge: bool = False, adapter_names: Optional[List[str]] = None) -> nn.Module:
        return self._unload_and_optionally_merge(True, safe_merge, adapter_names)

    def unload(self) -> nn.Module:
        return self._unload_and_optionally_merge(merge=False)

    def get_peft_config_as_dict(self, inference: bool = False):
        out = {}
        for name, cfg in self.peft_config.items():
            d = dict(cfg) if hasattr(cfg, "items") else dict(vars(cfg))
            if inference:
                d["inference_mode"] = True
            out[name] = d
        return out

    def lora_state_dict(self):
        """Only the adapter tensors (what a LoRA checkpoint stores)."""
        return {k: v for k, v in self.model.state_dict().items() if self.prefix in k}


# Same consolidation on the model side: ``LoraModel`` implements the whole tuner protocol (inject / enable / disable /
# merge / unload) that the reference declares in an abstract ``BaseTuner``.
BaseTuner = LoraModel
