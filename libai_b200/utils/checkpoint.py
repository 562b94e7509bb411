"""Checkpointing with the reference's on-disk layout and resume protocol.

Layout (reference docs Load_and_Save_Checkpoint.md:12-36, libai/utils/checkpoint.py:87-120,
:205-213, :342-378)::

    output_dir/
        last_checkpoint                      # basename of the newest checkpoint directory
        model_0000099/{model, optimizer, lr_scheduler, ...}
        model_final/…      model_best/…      # model_best is never written to last_checkpoint

Each entry is a ``torch.save`` file holding **unsharded logical tensors** under the model's own
parameter names (TP shards gathered, ZeRO partitions merged, all pipeline stages) written by
rank 0 — so a checkpoint can be resumed under a different (dp, tp, pp) layout.  OneFlow's binary
tensor format is not reproduced.
"""
from __future__ import annotations

import copy
import logging
import os
import re
import shutil
from collections import defaultdict
from typing import Any, Dict, List, NamedTuple, Optional, Tuple

import torch
from torch import nn

from libai_b200.parallel import state as pstate
from libai_b200.utils import distributed as dutil
from libai_b200.utils.file_io import PathManager


class _IncompatibleKeys(NamedTuple):
    missing_keys: List[str]
    unexpected_keys: List[str]
    incorrect_shapes: List[Tuple[str, Tuple[int, ...], Tuple[int, ...]]]


def _unwrap(model):
    return model.module if hasattr(model, "module") and isinstance(model.module, nn.Module) else model


class Checkpointer:
    """Save/load a model plus arbitrary ``state_dict()``/``load_state_dict()`` objects."""

    def __init__(self, model: nn.Module, save_dir: str = "", *, save_to_disk: bool = True, **checkpointables):
        self.model = _unwrap(model)
        self.checkpointables = copy.copy(checkpointables)
        self.logger = logging.getLogger(__name__)
        self.save_dir = save_dir
        self.save_to_disk = save_to_disk
        self.path_manager = PathManager

    # ------------------------------------------------------------------ save
    def save(self, name: str, **kwargs: Any) -> None:
        """Collective over all ranks: gathers logical tensors, rank 0 writes."""
        with self._params_materialized():
            payload = {"model": pstate.full_state_dict(self.model)}
        for key, obj in self.checkpointables.items():
            payload[key] = obj.state_dict()
        payload.update(kwargs)
        if dutil.is_main_process() and self.save_to_disk:
            target = os.path.join(self.save_dir, name)
            assert os.path.basename(target) == name, name
            os.makedirs(target, exist_ok=True)
            self.logger.info(f"Saving checkpoint to {target}")
            for key, obj in payload.items():
                if key == "iteration":
                    continue  # recovered from the directory name on load
                tmp = os.path.join(target, key + ".tmp")
                torch.save(obj, tmp)
                os.replace(tmp, os.path.join(target, key))
            if name != "model_best":
                self.tag_last_checkpoint(name)
        dutil.synchronize()

    # ------------------------------------------------------------------ load
    def load(self, path: str, checkpointables: Optional[List[str]] = None) -> Dict[str, Any]:
        if not path:
            self.logger.info("No checkpoint found. Training model from scratch")
            return {}
        self.logger.info(f"Loading checkpoint from {path}")
        ckpt = self._load_file(path)
        with self._params_materialized(writeback=True):
            incompatible = self._load_model(ckpt)
            if incompatible is not None:
                self._log_incompatible_keys(incompatible)
            # the optimizer's fp32 master weights were snapshotted from the pre-load parameters: re-derive them (an
            # optimizer state with ``master`` entries, loaded below, then overrides this with the exact fp32 values)
            for obj in self.checkpointables.values():
                if hasattr(obj, "refresh_master"):
                    obj.refresh_master()
        for key in list(self.checkpointables if checkpointables is None else checkpointables):
            if key in ckpt:
                self.logger.info(f"Loading {key} from {path}")
                self.checkpointables[key].load_state_dict(ckpt.pop(key))
        return ckpt

    def _params_materialized(self, writeback: bool = False):
        """ZeRO stage 3 keeps block parameters as 1/dp shards between uses: reading / writing the model's tensors needs
        them gathered (``FlatOptimizer.params_materialized``); a no-op otherwise."""
        from contextlib import nullcontext

        for obj in self.checkpointables.values():
            if hasattr(obj, "params_materialized"):
                return obj.params_materialized(writeback=writeback)
        return nullcontext()

    def has_checkpoint(self) -> bool:
        return os.path.exists(os.path.join(self.save_dir, "last_checkpoint"))

    def get_checkpoint_file(self) -> str:
        marker = os.path.join(self.save_dir, "last_checkpoint")
        try:
            with open(marker, "r") as f:
                last = f.read().strip()
        except IOError:
            return ""  # a concurrent writer may have removed it; treat as absent
        return os.path.join(self.save_dir, last)

    def resume_or_load(self, path: str, *, resume: bool = True) -> Dict[str, Any]:
        """``resume=True`` and a ``last_checkpoint`` exists → load it with all checkpointables;
        otherwise load *weights only* from ``path``."""
        if resume and self.has_checkpoint():
            return self.load(self.get_checkpoint_file())
        return self.load(path, checkpointables=[])

    def tag_last_checkpoint(self, last_filename_basename: str) -> None:
        marker = os.path.join(self.save_dir, "last_checkpoint")
        tmp = marker + ".tmp"
        with open(tmp, "w") as f:      # atomic: a kill mid-write must not leave an empty marker behind
            f.write(last_filename_basename)
            f.flush()
            os.fsync(f.fileno())
        os.replace(tmp, marker)

    # ------------------------------------------------------------------ internals
    def _load_file(self, path: str) -> Dict[str, Any]:
        """Read a checkpoint *directory* (one file per key) or a single ``torch.save`` file."""
        data: Dict[str, Any] = {}
        if os.path.isdir(path):
            for key in sorted(os.listdir(path)):
                fp = os.path.join(path, key)
                if os.path.isfile(fp) and not key.endswith(".tmp"):
                    data[key] = torch.load(fp, map_location="cpu", weights_only=False)
            m = re.search(r"_(\d+)$", os.path.basename(os.path.normpath(path)))
            if m:
                data["iter"] = int(m.group(1))
        else:
            obj = torch.load(path, map_location="cpu", weights_only=False)
            data = obj if isinstance(obj, dict) and "model" in obj else {"model": obj}
        return data

    def _load_model(self, checkpoint: Dict[str, Any]) -> _IncompatibleKeys:
        state = checkpoint.pop("model")
        _strip_prefix_if_present(state, "module.")
        missing, unexpected, bad = pstate.load_full_state_dict(self.model, state, strict=False)
        for k, got, want in bad:
            self.logger.warning(
                f"Skip loading parameter '{k}' to the model due to incompatible shapes: "
                f"{got} in the checkpoint but {want} in the model! You might want to double check "
                "if this is expected."
            )
        return _IncompatibleKeys(missing, unexpected, bad)

    def _log_incompatible_keys(self, inc: _IncompatibleKeys) -> None:
        if inc.missing_keys:
            self.logger.info(get_missing_parameters_message(inc.missing_keys))
        if inc.unexpected_keys:
            self.logger.info(get_unexpected_parameters_message(inc.unexpected_keys))


class PeriodicCheckpointer:
    """Save every ``period`` iterations, keep at most ``max_to_keep`` (rank 0 deletes), and write
    ``{prefix}_final`` at ``max_iter`` (spec: reference checkpoint.py:309-390)."""

    def __init__(
        self,
        checkpointer: Checkpointer,
        period: int,
        max_iter: Optional[int] = None,
        max_to_keep: Optional[int] = None,
        file_prefix: str = "model",
    ):
        self.checkpointer = checkpointer
        self.period = int(period)
        self.max_iter = max_iter
        if max_to_keep is not None:
            assert max_to_keep > 0
        self.max_to_keep = max_to_keep
        self.recent_checkpoints: List[str] = []
        self.file_prefix = file_prefix

    def step(self, iteration: int, **kwargs: Any) -> None:
        iteration = int(iteration)
        extra = {"iteration": iteration}
        extra.update(kwargs)
        if (iteration + 1) % self.period == 0:
            name = f"{self.file_prefix}_{iteration:07d}"
            self.checkpointer.save(name, **extra)
            if self.max_to_keep is not None:
                self.recent_checkpoints.append(os.path.join(self.checkpointer.save_dir, name))
                if len(self.recent_checkpoints) > self.max_to_keep:
                    victim = self.recent_checkpoints.pop(0)
                    if (
                        dutil.is_main_process()
                        and os.path.exists(victim)
                        and not victim.endswith(f"{self.file_prefix}_final")
                    ):
                        shutil.rmtree(victim, ignore_errors=True)
        if self.max_iter is not None and iteration >= self.max_iter - 1:
            self.checkpointer.save(f"{self.file_prefix}_final", **extra)

    def save(self, name: str, **kwargs: Any) -> None:
        self.checkpointer.save(name, **kwargs)


# ---- key-reporting helpers (pure string utilities) ------------------------------------------
def _strip_prefix_if_present(state_dict: Dict[str, Any], prefix: str) -> None:
    keys = sorted(state_dict.keys())
    if not keys or not all(k.startswith(prefix) for k in keys):
        return
    for k in keys:
        state_dict[k[len(prefix) :]] = state_dict.pop(k)


def _group_checkpoint_keys(keys: List[str]) -> Dict[str, List[str]]:
    groups = defaultdict(list)
    for key in keys:
        pos = key.rfind(".")
        head, tail = (key[:pos], [key[pos + 1 :]]) if pos >= 0 else (key, [])
        groups[head].extend(tail)
    return groups


def _group_to_str(group: List[str]) -> str:
    if len(group) == 0:
        return ""
    if len(group) == 1:
        return "." + group[0]
    return ".{" + ", ".join(group) + "}"


def get_missing_parameters_message(keys: List[str]) -> str:
    groups = _group_checkpoint_keys(keys)
    msg = "Some model parameters or buffers are not found in the checkpoint:\n"
    msg += "\n".join("  " + k + _group_to_str(v) for k, v in groups.items())
    return msg


def get_unexpected_parameters_message(keys: List[str]) -> str:
    groups = _group_checkpoint_keys(keys)
    msg = "The checkpoint state_dict contains keys that are not used by the model:\n"
    msg += "\n".join("  " + k + _group_to_str(v) for k, v in groups.items())
    return msg
