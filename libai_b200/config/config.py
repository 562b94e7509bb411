"""``configurable`` decorator, ``try_get_key`` and ``get_config``.

Spec: reference libai/config/config.py:32-119 (configurable), :171-180 (try_get_key),
:183-198 (get_config).  A class decorated with ``@configurable`` on ``__init__`` may be built
either with explicit arguments or with a config node as first argument (routed through the
class' ``from_config`` classmethod).
"""
from __future__ import annotations

import functools
import inspect
import os

from .dictconfig import DictConfig, OmegaConf
from .lazy import LazyConfig

__all__ = ["configurable", "try_get_key", "get_config"]


def _is_cfg_call(*args, **kwargs) -> bool:
    if args and isinstance(args[0], DictConfig):
        return True
    return isinstance(kwargs.get("cfg", None), DictConfig)


def _explicit_args(from_config, *args, **kwargs) -> dict:
    sig = inspect.signature(from_config)
    names = list(sig.parameters)
    if not names or names[0] != "cfg":
        who = from_config.__name__ if inspect.isfunction(from_config) else f"{from_config.__self__}.from_config"
        raise TypeError(f"{who} must take 'cfg' as the first argument!")
    takes_var = any(
        p.kind in (p.VAR_POSITIONAL, p.VAR_KEYWORD) for p in sig.parameters.values()
    )
    if takes_var:
        return from_config(*args, **kwargs)
    passthrough = {k: kwargs.pop(k) for k in list(kwargs) if k not in sig.parameters}
    out = from_config(*args, **kwargs)
    out.update(passthrough)
    return out


def configurable(init_func=None, *, from_config=None):
    """See module docstring. Usage 1: decorate ``__init__``; usage 2: ``@configurable(from_config=f)``."""
    if init_func is not None:
        assert (
            inspect.isfunction(init_func) and from_config is None and init_func.__name__ == "__init__"
        ), "Incorrect use of @configurable. Check API documentation for examples."

        @functools.wraps(init_func)
        def wrapped_init(self, *args, **kwargs):
            fc = getattr(type(self), "from_config", None)
            if fc is None:
                raise AttributeError("Class with @configurable must have a 'from_config' classmethod.")
            if not inspect.ismethod(fc):
                raise TypeError("Class with @configurable must have a 'from_config' classmethod.")
            if _is_cfg_call(*args, **kwargs):
                init_func(self, **_explicit_args(fc, *args, **kwargs))
            else:
                init_func(self, *args, **kwargs)

        return wrapped_init

    if from_config is None:
        return configurable
    assert inspect.isfunction(from_config), "from_config argument of configurable must be a function!"

    def decorate(fn):
        @functools.wraps(fn)
        def wrapped(*args, **kwargs):
            if _is_cfg_call(*args, **kwargs):
                return fn(**_explicit_args(from_config, *args, **kwargs))
            return fn(*args, **kwargs)

        wrapped.from_config = from_config
        return wrapped

    return decorate


def try_get_key(cfg, *keys, default=None):
    """Return the first existing dotted key among ``keys`` else ``default``."""
    sentinel = object()
    for k in keys:
        v = OmegaConf.select(cfg, k, default=sentinel)
        if v is not sentinel:
            return v
    return default


def _configs_root() -> str:
    here = os.path.dirname(os.path.abspath(__file__))
    packaged = os.path.join(here, "configs")
    if os.path.isdir(packaged):
        return packaged
    return os.path.join(os.path.dirname(os.path.dirname(here)), "configs")


def get_config(config_path: str):
    """Load a config shipped with the framework, e.g. ``get_config("common/models/bert.py")``."""
    path = os.path.join(_configs_root(), config_path)
    if not os.path.exists(path):
        raise RuntimeError(f"{config_path} not available in libai_b200 configs!")
    return LazyConfig.load(path)
