"""Self-contained config containers (no omegaconf / hydra dependency).

The reference keeps its configs in ``omegaconf.DictConfig`` objects created with
``flags={"allow_objects": True}`` (reference: libai/config/lazy.py:112-123) and manipulates
them through ``OmegaConf.select / update / merge / to_container``
(reference: libai/config/lazy.py:362-401, libai/config/config.py:171-180).
omegaconf is not available in the target image, so this module provides the subset of that
behaviour the framework (and user configs) rely on:

* ``DictConfig`` – ordered mapping with attribute access, arbitrary python objects as values,
  nested ``dict``/``list`` values are wrapped on assignment.
* ``ListConfig`` – list counterpart.
* ``OmegaConf``  – namespace with ``create/select/update/merge/to_container/is_config/...``.
"""
from __future__ import annotations

import copy
from collections import abc
from typing import Any, Iterable, Iterator, Optional

_MISSING = object()


class ConfigAttributeError(AttributeError, KeyError):
    """Raised for a missing key, both for ``cfg.key`` and ``cfg["key"]``."""

    def __str__(self):  # KeyError quotes its message; keep it readable
        return str(self.args[0]) if self.args else ""


def _wrap(value: Any, parent_flags: Optional[dict] = None) -> Any:
    """Wrap plain containers so that nested access keeps attribute semantics."""
    if isinstance(value, (DictConfig, ListConfig)):
        return value
    if isinstance(value, dict):
        return DictConfig(value, flags=parent_flags)
    if isinstance(value, (list, tuple)) and not hasattr(value, "_fields"):
        # tuples are stored as lists exactly like omegaconf does; namedtuples stay objects
        return ListConfig(value, flags=parent_flags)
    return value


class DictConfig(abc.MutableMapping):
    __slots__ = ("_content", "_flags")

    def __init__(self, content: Optional[abc.Mapping] = None, flags: Optional[dict] = None, **kw):
        object.__setattr__(self, "_content", {})
        object.__setattr__(self, "_flags", dict(flags) if flags else {"allow_objects": True})
        if content is not None:
            if isinstance(content, DictConfig):
                content = content._content
            for k, v in content.items():
                self._content[k] = _wrap(v, self._flags)
        for k, v in kw.items():
            self._content[k] = _wrap(v, self._flags)

    # ---- mapping protocol -------------------------------------------------------------
    def __getitem__(self, key):
        try:
            return self._content[key]
        except KeyError:
            raise ConfigAttributeError(f"Missing key {key!r} (available: {list(self._content)})")

    def __setitem__(self, key, value):
        self._content[key] = _wrap(value, self._flags)

    def __delitem__(self, key):
        try:
            del self._content[key]
        except KeyError:
            raise ConfigAttributeError(f"Missing key {key!r}")

    def __iter__(self) -> Iterator:
        return iter(self._content)

    def __len__(self) -> int:
        return len(self._content)

    def __contains__(self, key) -> bool:
        return key in self._content

    # ---- attribute protocol -----------------------------------------------------------
    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        try:
            return self._content[name]
        except KeyError:
            raise ConfigAttributeError(
                f"Missing key {name!r} (available: {list(self._content)})"
            )

    def __setattr__(self, name, value):
        self._content[name] = _wrap(value, self._flags)

    def __delattr__(self, name):
        self.__delitem__(name)

    def __dir__(self):
        return list(self._content.keys())

    # ---- dict helpers -----------------------------------------------------------------
    def get(self, key, default=None):
        return self._content.get(key, default)

    def pop(self, key, default=_MISSING):
        if default is _MISSING:
            try:
                return self._content.pop(key)
            except KeyError:
                raise ConfigAttributeError(f"Missing key {key!r}")
        return self._content.pop(key, default)

    def keys(self):
        return self._content.keys()

    def values(self):
        return self._content.values()

    def items(self):
        return self._content.items()

    def update(self, other=(), **kw):  # shallow like dict.update (reference configs rely on it)
        if isinstance(other, abc.Mapping):
            for k in other:
                self[k] = other[k]
        else:
            for k, v in other:
                self[k] = v
        for k, v in kw.items():
            self[k] = v

    def setdefault(self, key, default=None):
        if key not in self._content:
            self[key] = default
        return self._content[key]

    def copy(self):
        return DictConfig(dict(self._content), flags=self._flags)

    # ---- misc -------------------------------------------------------------------------
    def __eq__(self, other):
        if isinstance(other, DictConfig):
            return self._content == other._content
        if isinstance(other, abc.Mapping):
            return to_container(self) == to_container(other)
        return NotImplemented

    def __ne__(self, other):
        r = self.__eq__(other)
        return r if r is NotImplemented else not r

    __hash__ = None

    def __repr__(self):
        return repr(to_container(self))

    def __deepcopy__(self, memo):
        new = DictConfig(flags=self._flags)
        memo[id(self)] = new
        for k, v in self._content.items():
            new._content[k] = copy.deepcopy(v, memo)
        return new

    def __copy__(self):
        return self.copy()

    def __getstate__(self):
        return {"content": self._content, "flags": self._flags}

    def __setstate__(self, state):
        object.__setattr__(self, "_content", state["content"])
        object.__setattr__(self, "_flags", state["flags"])

    def __reduce__(self):
        return (_rebuild_dict, (self._content, self._flags))


def _rebuild_dict(content, flags):
    d = DictConfig(flags=flags)
    d._content.update(content)
    return d


class ListConfig(abc.MutableSequence):
    __slots__ = ("_content", "_flags")

    def __init__(self, content: Optional[Iterable] = None, flags: Optional[dict] = None):
        self._flags = dict(flags) if flags else {"allow_objects": True}
        self._content = [_wrap(v, self._flags) for v in (content or [])]

    def __getitem__(self, idx):
        if isinstance(idx, slice):
            return ListConfig(self._content[idx], flags=self._flags)
        return self._content[idx]

    def __setitem__(self, idx, value):
        if isinstance(idx, slice):
            self._content[idx] = [_wrap(v, self._flags) for v in value]
        else:
            self._content[idx] = _wrap(value, self._flags)

    def __delitem__(self, idx):
        del self._content[idx]

    def __len__(self):
        return len(self._content)

    def insert(self, idx, value):
        self._content.insert(idx, _wrap(value, self._flags))

    def _iter_ex(self, resolve=True):
        return iter(self._content)

    def __eq__(self, other):
        if isinstance(other, ListConfig):
            return self._content == other._content
        if isinstance(other, (list, tuple)):
            return to_container(self) == list(other)
        return NotImplemented

    __hash__ = None

    def __add__(self, other):
        return ListConfig(list(self._content) + list(other), flags=self._flags)

    def __radd__(self, other):
        return ListConfig(list(other) + list(self._content), flags=self._flags)

    def __repr__(self):
        return repr(to_container(self))

    def __deepcopy__(self, memo):
        new = ListConfig(flags=self._flags)
        memo[id(self)] = new
        new._content = [copy.deepcopy(v, memo) for v in self._content]
        return new

    def __reduce__(self):
        return (_rebuild_list, (self._content, self._flags))


def _rebuild_list(content, flags):
    lst = ListConfig(flags=flags)
    lst._content = list(content)
    return lst


# --------------------------------------------------------------------------------------
# OmegaConf-like helpers
# --------------------------------------------------------------------------------------
def to_container(cfg: Any, resolve: bool = True) -> Any:
    """Recursively convert config containers to plain ``dict`` / ``list``."""
    if isinstance(cfg, (DictConfig, abc.Mapping)):
        return {k: to_container(v, resolve) for k, v in cfg.items()}
    if isinstance(cfg, (ListConfig, list)):
        return [to_container(v, resolve) for v in cfg]
    return cfg


def is_dict(obj) -> bool:
    return isinstance(obj, DictConfig)


def is_list(obj) -> bool:
    return isinstance(obj, ListConfig)


def is_config(obj) -> bool:
    return isinstance(obj, (DictConfig, ListConfig))


def create(obj: Any = None, flags: Optional[dict] = None):
    if obj is None:
        return DictConfig(flags=flags)
    if isinstance(obj, str):
        import yaml

        obj = yaml.safe_load(obj)
    if isinstance(obj, (DictConfig, abc.Mapping)):
        return DictConfig(obj, flags=flags)
    if isinstance(obj, (ListConfig, list, tuple)):
        return ListConfig(obj, flags=flags)
    raise TypeError(f"Cannot create a config from {type(obj)}")


def _step(node, part):
    if isinstance(node, (DictConfig, abc.Mapping)):
        if part in node:
            return node[part]
        return _MISSING
    if isinstance(node, (ListConfig, list)):
        try:
            return node[int(part)]
        except (ValueError, IndexError):
            return _MISSING
    return _MISSING


def select(cfg, key: str, default: Any = None):
    """``select(cfg, "a.b.0.c")``; returns ``default`` when any hop is absent."""
    node = cfg
    if key in ("", None):
        return node
    for part in str(key).split("."):
        node = _step(node, part)
        if node is _MISSING:
            return default
    return node


def merge(*cfgs):
    """Deep-merge mappings left to right into a new DictConfig (lists are replaced)."""
    out = DictConfig()
    for c in cfgs:
        if c is None:
            continue
        _merge_into(out, c)
    return out


def _merge_into(dst: DictConfig, src) -> None:
    for k, v in src.items():
        if k in dst and isinstance(dst[k], DictConfig) and isinstance(v, (DictConfig, abc.Mapping)):
            _merge_into(dst[k], v)
        else:
            dst[k] = copy.deepcopy(v) if isinstance(v, (DictConfig, ListConfig)) else v


def update(cfg, key: str, value: Any, merge: bool = True, force_add: bool = True) -> None:
    """Set ``cfg.<dotted key> = value`` creating intermediate dicts; dict values are merged."""
    parts = str(key).split(".")
    node = cfg
    for i, part in enumerate(parts[:-1]):
        nxt = _step(node, part)
        if nxt is _MISSING or nxt is None:
            nxt = DictConfig()
            if isinstance(node, (ListConfig, list)):
                node[int(part)] = nxt
            else:
                node[part] = nxt
            nxt = _step(node, part)
        if not isinstance(nxt, (DictConfig, ListConfig)):
            raise KeyError(
                f"Trying to update key {key}, but {'.'.join(parts[: i + 1])} "
                f"is not a config, but has type {type(nxt)}."
            )
        node = nxt
    last = parts[-1]
    if isinstance(node, (ListConfig, list)):
        node[int(last)] = value
        return
    cur = node.get(last, _MISSING) if isinstance(node, DictConfig) else _MISSING
    if merge and isinstance(cur, DictConfig) and isinstance(value, (DictConfig, abc.Mapping)):
        _merge_into(cur, value)
    else:
        node[last] = value


class OmegaConf:
    """Namespace mirroring the handful of ``omegaconf.OmegaConf`` entry points that are used."""

    create = staticmethod(create)
    select = staticmethod(select)
    update = staticmethod(update)
    merge = staticmethod(merge)
    to_container = staticmethod(to_container)
    is_config = staticmethod(is_config)
    is_dict = staticmethod(is_dict)
    is_list = staticmethod(is_list)

    @staticmethod
    def to_yaml(cfg) -> str:
        import yaml

        return yaml.dump(to_container(cfg), default_flow_style=None, allow_unicode=True, width=9999)

    @staticmethod
    def load(path: str):
        import yaml

        with open(path, "r", encoding="utf-8") as f:
            return create(yaml.unsafe_load(f))

    @staticmethod
    def save(cfg, path: str):
        with open(path, "w", encoding="utf-8") as f:
            f.write(OmegaConf.to_yaml(cfg))
