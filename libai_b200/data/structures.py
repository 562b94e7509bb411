"""Batch containers.

Spec: reference libai/data/structures.py:25-196 — ``Instance`` (ordered named fields, ``stack``
collate) and ``DistTensorData`` (a tensor + how it is distributed: ``sbp_list`` default
``["split_0", "broadcast"]`` = batch-sharded over DP / replicated over TP, and ``placement_idx``
0 → first pipeline stage, −1 → last stage, used for labels).

With explicit process groups "making a tensor global" reduces to: the DP shard is already selected
by the sampler, TP ranks of one DP replica read identical samples, so ``to_global`` only has to
move the local tensor to the device (pinned-memory H2D, non-blocking).  ``placement_idx`` is kept
so the pipeline engine knows which stage consumes a field (the last stage reads its labels from
its own loader copy – same sampler indices – instead of receiving them over p2p).
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Any, List

import torch

from libai_b200.utils import distributed as dutil


@dataclass
class DistTensorData:
    tensor: torch.Tensor
    sbp_list: list = field(default_factory=lambda: ["split_0", "broadcast"])
    placement_idx: int = 0

    def to_global(self, sbp=None, placement=None, device_type=None):
        """Move to this rank's device (name kept from the reference API)."""
        dev = dutil.get_device() if device_type in (None, "cuda") else torch.device(device_type)
        if dutil.get_dist_util().device_type == "cpu":
            dev = torch.device("cpu")
        t = self.tensor
        if dev.type == "cuda" and t.device.type == "cpu":
            if not t.is_pinned():
                t = t.pin_memory()
            t = t.to(dev, non_blocking=True)
        else:
            t = t.to(dev)
        self.tensor = t
        return self

    def needed_on_this_stage(self) -> bool:
        topo = dutil.get_dist_util()
        if topo.pipeline_parallel_size == 1:
            return True
        return topo.get_layer_stage_id(self.placement_idx) == topo.pp_rank

    @staticmethod
    def stack(items: List["DistTensorData"]) -> "DistTensorData":
        if not isinstance(items[0].tensor, torch.Tensor):
            raise TypeError(
                "DistTensorData.tensor must be a torch.Tensor, but got {}. "
                "Please check the return values of `__getitem__` in dataset.".format(type(items[0].tensor))
            )
        assert len(items) > 0
        first = items[0]
        if len(items) == 1:
            first.tensor = first.tensor.unsqueeze(0)
            return first
        for d in items:
            assert d.tensor.size() == first.tensor.size(), (
                f"tensor shape is not equal, {d.tensor.size()} != {first.tensor.size()}"
            )
            assert d.sbp_list == first.sbp_list, f"sbp_list is not equal, {d.sbp_list} != {first.sbp_list}!"
            assert d.placement_idx == first.placement_idx, (
                f"placement_idx is not equal, {d.placement_idx} != {first.placement_idx}"
            )
        return DistTensorData(
            torch.stack([d.tensor for d in items], dim=0), sbp_list=first.sbp_list, placement_idx=first.placement_idx
        )


class Instance:
    """A sample (or batch) as ordered named fields: ``Instance(input_ids=DistTensorData(...), ...)``.
    Field names equal the model's ``forward`` keyword names."""

    def __init__(self, **kwargs):
        self._fields = OrderedDict()
        for k, v in kwargs.items():
            self.set(k, v)

    def __setattr__(self, name: str, val: Any) -> None:
        if name.startswith("_"):
            super().__setattr__(name, val)
        else:
            self.set(name, val)

    def __getattr__(self, name: str):
        if name == "_fields" or name not in self._fields:
            raise AttributeError(f"Cannot find field '{name}' in the given Instance!")
        return self._fields[name]

    def set(self, name: str, value: Any):
        self._fields[name] = value

    def has(self, name: str):
        return name in self._fields

    def remove(self, name: str):
        del self._fields[name]

    def get(self, name: str):
        return self._fields[name]

    def get_fields(self):
        return self._fields

    def __len__(self):
        return len(self._fields.keys())

    def __iter__(self):
        raise NotImplementedError("`Instances` object is not iterable!")

    @staticmethod
    def stack(instance_lists: List["Instance"]) -> "Instance":
        assert all(isinstance(i, Instance) for i in instance_lists)
        assert len(instance_lists) > 0
        out = Instance()
        for k in instance_lists[0]._fields.keys():
            vals = [i.get(k) for i in instance_lists]
            v0 = vals[0]
            if isinstance(v0, torch.Tensor):
                vals = torch.stack(vals, dim=0)
            elif isinstance(v0, list):
                pass
            elif hasattr(type(v0), "stack"):
                vals = type(v0).stack(vals)
            else:
                raise ValueError("Unsupported type {} for stack.".format(type(v0)))
            out.set(k, vals)
        return out

    def __str__(self):
        body = ", ".join(f"{k}: {v}" for k, v in self._fields.items())
        return f"{self.__class__.__name__}(fields=[{body}])"

    __repr__ = __str__
