"""Build python objects from ``{_target_: ..., **kwargs}`` records.

Spec: reference libai/config/instantiate.py:130-201 (``instantiate``/``instantiate_cfg``),
:61-82 (``dump_dataclass``).  ``_target_`` may be a callable, a dotted string, or itself a lazy
record; ``_recursive_: False`` stops descent below the node that carries it; list configs are
instantiated element-wise and come back as ``ListConfig``; plain lists come back as lists.
"""
from __future__ import annotations

import dataclasses
from collections import abc
from typing import Any

from .dictconfig import DictConfig, ListConfig, OmegaConf
from .lazy import _convert_target_to_string, locate

__all__ = ["dump_dataclass", "instantiate", "InstantiationException"]

_TARGET = "_target_"
_RECURSIVE = "_recursive_"


class InstantiationException(Exception):
    pass


def dump_dataclass(obj: Any):
    """Recursively turn a dataclass *instance* into an instantiable dict."""
    assert dataclasses.is_dataclass(obj) and not isinstance(
        obj, type
    ), "dump_dataclass() requires an instance of a dataclass."
    out = {_TARGET: _convert_target_to_string(type(obj))}
    for field in dataclasses.fields(obj):
        val = getattr(obj, field.name)
        if dataclasses.is_dataclass(val):
            val = dump_dataclass(val)
        elif isinstance(val, (list, tuple)):
            val = [dump_dataclass(x) if dataclasses.is_dataclass(x) else x for x in val]
        out[field.name] = val
    return out


def _has_target(x) -> bool:
    return isinstance(x, (DictConfig, abc.Mapping)) and _TARGET in x


def _resolve_callable(target):
    if isinstance(target, str):
        try:
            target = locate(target)
        except Exception as e:
            raise InstantiationException(
                f"Error locating target '{target}', see chained exception above."
            ) from e
    if not callable(target):
        raise InstantiationException(
            f"Expected a callable target, got '{target}' of type '{type(target).__name__}'"
        )
    return target


def _build(node, recursive: bool):
    if node is None:
        return None
    if isinstance(node, (DictConfig, abc.Mapping)):
        if _RECURSIVE in node:
            recursive = node[_RECURSIVE]
        if not isinstance(recursive, bool):
            raise TypeError(f"Instantiation: _recursive_ flag must be a bool, got {type(recursive)}")
        if not _has_target(node):
            return node
        fn = _resolve_callable(instantiate(node.get(_TARGET)))
        kwargs = {}
        for k, v in node.items():
            if k in (_TARGET, _RECURSIVE):
                continue
            kwargs[k] = _build(v, recursive) if recursive else v
        try:
            return fn(**kwargs)
        except Exception as e:
            name = _convert_target_to_string(fn) if hasattr(fn, "__qualname__") else repr(fn)
            raise InstantiationException(f"Error in call to target '{name}':\n{e!r}") from e
    if isinstance(node, ListConfig):
        return ListConfig([_build(x, recursive) for x in node])
    if isinstance(node, list):
        return [_build(x, recursive) for x in node]
    return node


def instantiate_cfg(cfg: Any, recursive: bool = True):
    """Instantiate an already-prepared config node (no keyword merging) — the second half of :func:`instantiate`,
    public because the reference exposes it (libai/config/instantiate.py:164-201)."""
    return _build(cfg, recursive)


def instantiate(cfg, **kwargs: Any) -> Any:
    """Recursively instantiate ``cfg``; extra ``kwargs`` are merged over the top-level record."""
    if cfg is None:
        return None
    recursive = kwargs.pop(_RECURSIVE, True)
    if isinstance(cfg, (DictConfig, abc.Mapping)):
        if kwargs:
            cfg = OmegaConf.merge(cfg, kwargs)
        return _build(cfg, recursive)
    if isinstance(cfg, (ListConfig, list)):
        return _build(cfg, recursive)
    return cfg
