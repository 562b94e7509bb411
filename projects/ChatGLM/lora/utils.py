"""Target matching and sub-module lookup (reference projects/ChatGLM/lora/utils.py)."""
import re
from typing import Optional, Union


def check_target_module_exists(config, key: str) -> Union[bool, Optional[re.Match]]:
    """``target_modules`` is a regex (full match) or a list of name suffixes; ``layers_to_transform`` /
    ``layers_pattern`` optionally restrict the layer indices."""
    targets = config.get("target_modules") if hasattr(config, "get") else config.target_modules
    if isinstance(targets, str):
        found = re.fullmatch(targets, key)
    else:
        found = key in targets or any(key.endswith(f".{t}") for t in targets)
        layers = config.get("layers_to_transform", None) if hasattr(config, "get") else None
        if found and layers is not None:
            pattern = config.get("layers_pattern", None) or "layers"
            m = re.match(rf".*\.{pattern}\.(\d+)\.", key)
            idx = int(m.group(1)) if m else None
            layers = [layers] if isinstance(layers, int) else list(layers)
            found = idx in layers
    return found


def _get_submodules(model, key):
    parent = model.get_submodule(".".join(key.split(".")[:-1]))
    target_name = key.split(".")[-1]
    return parent, model.get_submodule(key), target_name
