"""Python-file ("lazy") configs: ``LazyCall`` + ``LazyConfig``.

Behavioural spec taken from the reference (libai/config/lazy.py:89-123 ``LazyCall``,
:168-224 relative-import patching, :249-301 ``load``, :304-359 ``save``, :362-401
``apply_overrides``, :404-463 ``to_py``).  The implementation is independent: configs are held
in the in-repo :mod:`dictconfig` containers, CLI overrides are parsed by :mod:`overrides`, and
config files are executed by a small loader object (`_ConfigFileLoader`) that resolves relative
imports by *file location* (no packages/``__init__`` needed, nothing cached in ``sys.modules``).
"""
from __future__ import annotations

import ast
import builtins
import importlib
import inspect
import logging
import os
import pydoc
import threading
import types
import uuid
from collections import abc
from contextlib import contextmanager
from copy import deepcopy
from dataclasses import is_dataclass
from typing import Any, List, Tuple, Union

import yaml

from .dictconfig import DictConfig, ListConfig, OmegaConf
from .overrides import OverridesParser

__all__ = ["LazyCall", "LazyConfig", "locate"]

_ALLOW_OBJECTS = {"allow_objects": True}


# --------------------------------------------------------------------------------------
# dotted-name <-> object
# --------------------------------------------------------------------------------------
def locate(name: str) -> Any:
    """Resolve ``"pkg.mod.Class.attr"`` to the object; raise ``ImportError`` if impossible."""
    obj = pydoc.locate(name)
    if obj is not None:
        return obj
    # pydoc gives up on some layouts (attribute of a class inside a module alias ...): walk manually
    parts = name.split(".")
    for split in range(len(parts), 0, -1):
        mod_name = ".".join(parts[:split])
        try:
            obj = importlib.import_module(mod_name)
        except Exception:
            continue
        try:
            for attr in parts[split:]:
                obj = getattr(obj, attr)
        except AttributeError:
            continue
        return obj
    raise ImportError(f"Cannot dynamically locate object {name}!")


def _convert_target_to_string(t: Any) -> str:
    """Shortest importable dotted name for ``t`` (inverse of :func:`locate`)."""
    module, qualname = t.__module__, t.__qualname__
    pieces = module.split(".")
    for k in range(1, len(pieces)):
        candidate = ".".join(pieces[:k]) + "." + qualname
        try:
            if locate(candidate) is t:
                return candidate
        except ImportError:
            pass
    return f"{module}.{qualname}"


# --------------------------------------------------------------------------------------
# LazyCall
# --------------------------------------------------------------------------------------
class LazyCall:
    """``LazyCall(fn)(**kwargs)`` records a call as ``DictConfig{_target_: fn, **kwargs}``.

    The record is built later by :func:`libai_b200.config.instantiate`.  Only keyword arguments
    are accepted (same contract as the reference).
    """

    def __init__(self, target):
        if not (callable(target) or isinstance(target, (str, abc.Mapping))):
            raise TypeError(
                f"target of LazyCall must be a callable or defines a callable! Got {target}"
            )
        self._target = target

    def __call__(self, **kwargs):
        target = self._target
        if is_dataclass(target):
            # keep dataclass *types* out of the container: store their dotted name
            target = _convert_target_to_string(target)
        kwargs["_target_"] = target
        return DictConfig(content=kwargs, flags=_ALLOW_OBJECTS)


def _visit_dict_config(cfg, func):
    if isinstance(cfg, DictConfig):
        func(cfg)
        for v in cfg.values():
            _visit_dict_config(v, func)
    elif isinstance(cfg, ListConfig):
        for v in cfg:
            _visit_dict_config(v, func)


def _cast_to_config(obj):
    if isinstance(obj, dict):
        return DictConfig(obj, flags=_ALLOW_OBJECTS)
    return obj


# --------------------------------------------------------------------------------------
# config-file execution with location-based relative imports
# --------------------------------------------------------------------------------------
_CFG_PACKAGE_NAME = "libai_b200._cfg_loader"
_import_lock = threading.RLock()


def _pseudo_package(filename: str) -> str:
    return f"{_CFG_PACKAGE_NAME}{uuid.uuid4().hex[:6]}.{os.path.basename(filename)}"


def _read_and_check(filename: str) -> str:
    with open(filename, "r", encoding="utf-8") as f:
        src = f.read()
    try:
        ast.parse(src)
    except SyntaxError as e:
        raise SyntaxError(f"Config file {filename} has syntax error!") from e
    return src


class _ConfigFileLoader:
    """Executes config files; relative ``from .x import y`` are resolved on disk."""

    def resolve(self, importer_file: str, rel_name: str, level: int) -> str:
        base = os.path.dirname(importer_file)
        for _ in range(level - 1):
            base = os.path.dirname(base)
        path = os.path.join(base, *[p for p in rel_name.split(".") if p])
        if not path.endswith(".py"):
            path += ".py"
        if not os.path.isfile(path):
            raise ImportError(
                f"Cannot import name {rel_name} from {importer_file}: {path} has to exist."
            )
        return path

    def exec_as_module(self, path: str, fromlist=()) -> types.ModuleType:
        src = _read_and_check(path)
        mod = types.ModuleType(_pseudo_package(path))
        mod.__file__ = path
        mod.__package__ = mod.__name__
        exec(compile(src, path, "exec"), mod.__dict__)
        for name in fromlist or ():
            if name in mod.__dict__:  # imported dicts become DictConfig automatically
                mod.__dict__[name] = _cast_to_config(mod.__dict__[name])
        return mod

    def exec_top(self, path: str) -> dict:
        src = _read_and_check(path)
        ns = {"__file__": path, "__package__": _pseudo_package(path)}
        exec(compile(src, path, "exec"), ns)
        return ns


_LOADER = _ConfigFileLoader()


@contextmanager
def _patch_import():
    """Temporarily route relative imports issued *from config files* through `_LOADER`."""
    with _import_lock:
        original = builtins.__import__

        def hooked(name, globals=None, locals=None, fromlist=(), level=0):
            pkg = (globals or {}).get("__package__", "") or ""
            if level != 0 and globals is not None and pkg.startswith(_CFG_PACKAGE_NAME):
                target = _LOADER.resolve(globals["__file__"], name, level)
                return _LOADER.exec_as_module(target, fromlist)
            return original(name, globals, locals, fromlist=fromlist, level=level)

        builtins.__import__ = hooked
        try:
            yield hooked
        finally:
            builtins.__import__ = original


# --------------------------------------------------------------------------------------
# LazyConfig
# --------------------------------------------------------------------------------------
class LazyConfig:
    """Load / save / override configs that may contain lazily-constructed objects."""

    @staticmethod
    def load_rel(filename: str, keys: Union[None, str, Tuple[str, ...]] = None):
        """Like :meth:`load` but ``filename`` is relative to the *caller's* source file."""
        caller = inspect.stack()[1]
        caller_file = caller[0].f_code.co_filename
        assert caller_file != "<string>", "load_rel Unable to find caller"
        return LazyConfig.load(os.path.join(os.path.dirname(caller_file), filename), keys)

    @staticmethod
    def load(filename: str, keys: Union[None, str, Tuple[str, ...]] = None):
        filename = filename.replace("/./", "/")
        ext = os.path.splitext(filename)[1]
        if ext not in (".py", ".yaml", ".yml"):
            raise ValueError(f"Config file {filename} has to be a python or yaml file.")
        if ext == ".py":
            with _patch_import():
                content = _LOADER.exec_top(filename)
        else:
            with open(filename, "r", encoding="utf-8") as f:
                content = OmegaConf.create(yaml.unsafe_load(f), flags=_ALLOW_OBJECTS)

        if keys is not None:
            if isinstance(keys, str):
                return _cast_to_config(content[keys])
            return tuple(_cast_to_config(content[k]) for k in keys)
        if ext == ".py":
            content = DictConfig(
                {
                    k: _cast_to_config(v)
                    for k, v in content.items()
                    if isinstance(v, (DictConfig, ListConfig, dict)) and not k.startswith("_")
                },
                flags=_ALLOW_OBJECTS,
            )
        return content

    @staticmethod
    def save(cfg, filename: str):
        """Dump to YAML; if the config holds un-dumpable objects also write ``<filename>.pkl``."""
        logger = logging.getLogger(__name__)
        try:
            cfg = deepcopy(cfg)
        except Exception:
            pass
        else:

            def _stringify_target(node):
                if "_target_" in node and callable(node["_target_"]):
                    try:
                        node["_target_"] = _convert_target_to_string(node["_target_"])
                    except AttributeError:
                        pass

            _visit_dict_config(cfg, _stringify_target)

        need_pickle = False
        try:
            plain = OmegaConf.to_container(cfg, resolve=False)
            text = yaml.dump(plain, default_flow_style=None, allow_unicode=True, width=9999)
            with open(filename, "w", encoding="utf-8") as f:
                f.write(text)
            try:
                yaml.unsafe_load(text)
            except Exception:
                logger.warning(
                    "The config contains objects that cannot serialize to a valid yaml. "
                    f"{filename} is human-readable but cannot be loaded."
                )
                need_pickle = True
        except Exception:
            logger.exception("Unable to serialize the config to yaml. Error:")
            need_pickle = True

        if need_pickle:
            import cloudpickle

            try:
                with open(filename + ".pkl", "wb") as f:
                    cloudpickle.dump(cfg, f)
                logger.warning(f"Config is saved using cloudpickle at {filename}.pkl.")
            except Exception:
                pass

    @staticmethod
    def apply_overrides(cfg, overrides: List[str]):
        """In-place ``a.b=c`` overrides; a non-config intermediate node raises ``KeyError``."""

        def checked_update(key, value):
            parts = key.split(".")
            for n in range(1, len(parts)):
                prefix = ".".join(parts[:n])
                node = OmegaConf.select(cfg, prefix, default=None)
                if node is None:
                    break
                if not OmegaConf.is_config(node):
                    raise KeyError(
                        f"Trying to update key {key}, but {prefix} "
                        f"is not a config, but has type {type(node)}."
                    )
            OmegaConf.update(cfg, key, value, merge=True)

        for o in OverridesParser.create().parse_overrides(list(overrides)):
            if o.is_delete():
                raise NotImplementedError("deletion is not yet a supported override")
            checked_update(o.key_or_group, o.value())
        return cfg

    @staticmethod
    def to_py(cfg, prefix: str = "cfg."):
        """Render a config as python-like pseudo code (one top-level assignment per line)."""
        plain = OmegaConf.to_container(cfg, resolve=True)

        def render(obj, lead=None, in_call=False):
            lead = lead or []
            if isinstance(obj, abc.Mapping) and "_target_" in obj:
                obj = dict(obj)
                tgt = obj.pop("_target_")
                tgt = tgt if isinstance(tgt, str) else _convert_target_to_string(tgt)
                args = ", ".join(f"{k}={render(v, in_call=True)}" for k, v in sorted(obj.items()))
                return "".join(lead) + f"{tgt}({args})"
            if isinstance(obj, abc.Mapping) and not in_call:
                lines = []
                for k, v in sorted(obj.items()):
                    if isinstance(v, abc.Mapping) and "_target_" not in v:
                        lines.append(render(v, lead=lead + [k + "."]))
                    else:
                        lines.append(f"{''.join(lead)}{k}={render(v)}")
                return "\n".join(lines)
            if isinstance(obj, abc.Mapping):
                body = ",".join(f"{k!r}: {render(v, in_call=in_call)}" for k, v in sorted(obj.items()))
                return "{" + body + "}"
            if isinstance(obj, list):
                return "[" + ",".join(render(x, in_call=True) for x in obj) + "]"
            return repr(obj)

        text = render(plain, lead=[prefix])
        try:
            import black

            return black.format_str(text, mode=black.Mode())
        except Exception:
            return _mini_format(text)


def _mini_format(text: str) -> str:
    """Fallback formatter when ``black`` is absent: normalise spacing around top-level ``=``."""
    out = []
    for line in text.splitlines():
        depth, cut = 0, None
        for i, ch in enumerate(line):
            if ch in "([{":
                depth += 1
            elif ch in ")]}":
                depth -= 1
            elif ch == "=" and depth == 0:
                cut = i
                break
        if cut is None:
            out.append(line)
            continue
        lhs, rhs = line[:cut], line[cut + 1 :]
        try:
            rhs = ast.unparse(ast.parse(rhs, mode="eval")).replace("'", '"')
        except Exception:
            pass
        out.append(f"{lhs} = {rhs}")
    return "\n".join(out) + "\n"
