"""Logical (unsharded) ⇄ per-rank (TP-sharded, PP-placed) state dictionaries.

The reference stores *global* tensors gathered to rank 0 (``flow.save(..., global_dst_rank=0)``,
reference libai/utils/checkpoint.py:87-120) which makes checkpoints parallelism-agnostic, and
re-shards on load (:302-306).  With explicit process groups the same contract is implemented
here: every parameter carries ``tp_dim`` (the dimension split across the tensor-parallel group,
or ``None`` when replicated); pipeline stages own disjoint parameter subsets (non-owned
parameters live on the ``meta`` device).
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Iterable, Optional

import torch
import torch.distributed as dist

from libai_b200.utils import distributed as dutil


def mark_tp(param: torch.Tensor, tp_dim: Optional[int], stride: int = 1) -> torch.Tensor:
    """Annotate a parameter with its tensor-parallel split dimension."""
    param.tp_dim = tp_dim
    param.tp_stride = stride
    return param


def tp_dim_of(t) -> Optional[int]:
    return getattr(t, "tp_dim", None)


def is_materialized(t: torch.Tensor) -> bool:
    return t.device.type != "meta"


def _named_tensors(module: torch.nn.Module):
    """(name, tensor) for parameters and persistent buffers, in ``state_dict`` order."""
    sd_keys = list(module.state_dict(keep_vars=True).items())
    return sd_keys


def gather_tp(t: torch.Tensor, tp_dim: Optional[int]) -> torch.Tensor:
    """Return the full logical tensor of a TP shard (identity when replicated / tp == 1)."""
    topo = dutil.get_dist_util()
    if tp_dim is None or topo.tensor_parallel_size == 1 or topo.tp_group is None:
        return t
    parts = [torch.empty_like(t) for _ in range(topo.tensor_parallel_size)]
    dist.all_gather(parts, t.contiguous(), group=topo.tp_group)
    return torch.cat(parts, dim=tp_dim)


def shard_tp(full: torch.Tensor, tp_dim: Optional[int]) -> torch.Tensor:
    """Slice this rank's TP shard out of a full logical tensor."""
    topo = dutil.get_dist_util()
    if tp_dim is None or topo.tensor_parallel_size == 1:
        return full
    n = full.shape[tp_dim]
    assert n % topo.tensor_parallel_size == 0, (
        f"dim {tp_dim} of size {n} is not divisible by tp={topo.tensor_parallel_size}"
    )
    per = n // topo.tensor_parallel_size
    return full.narrow(tp_dim, topo.tp_rank * per, per)


def full_state_dict(module: torch.nn.Module, to_cpu: bool = True) -> Dict[str, torch.Tensor]:
    """Collective: build the unsharded, all-stages state dict (complete on rank 0).

    Every rank must call it.  TP shards are all-gathered inside each TP group; pipeline stages
    then hand their tensors to rank 0 through the host (gloo) object channel.
    """
    topo = dutil.get_dist_util()
    local = OrderedDict()
    for name, t in _named_tensors(module):
        if not is_materialized(t):
            continue
        full = gather_tp(t.detach(), tp_dim_of(t))
        if topo.dp_rank == 0 and topo.tp_rank == 0:
            local[name] = full.cpu() if to_cpu else full
    if topo.pipeline_parallel_size == 1 or not dist.is_initialized():
        return local
    # merge the stage leaders' dictionaries on rank 0 (stage order == key order of the model)
    gathered = dutil.all_gather_py_object(local if (topo.dp_rank == 0 and topo.tp_rank == 0) else None)
    if dutil.get_rank() != 0:
        return local
    merged = OrderedDict()
    order = [k for k, _ in _named_tensors(module)]
    pool = {}
    for part in gathered:
        if part:
            pool.update(part)
    for k in order:
        if k in pool:
            merged[k] = pool[k]
    return merged


def load_full_state_dict(module: torch.nn.Module, state: Dict[str, torch.Tensor], strict: bool = False):
    """Copy this rank's shard of every tensor in ``state`` (logical, unsharded) into ``module``.

    Returns ``(missing_keys, unexpected_keys, mismatched)`` with the semantics of
    ``nn.Module.load_state_dict(strict=False)``; keys whose *logical* shape disagrees are dropped
    with a record in ``mismatched`` (reference behaviour: checkpoint.py:250-271).
    """
    missing, mismatched = [], []
    own = dict(_named_tensors(module))
    unexpected = [k for k in state if k not in own]
    topo = dutil.get_dist_util()
    with torch.no_grad():
        for name, t in own.items():
            if not is_materialized(t):
                continue  # belongs to another pipeline stage
            key = name
            if key not in state and getattr(t, "shared_from", None) is not None:
                # last-stage copy of a tied weight (pipeline parallelism): a checkpoint written with pp == 1, or an
                # HF-converted state dict, only carries the source tensor — fill the copy from it so both stay in sync
                prefix = name.rsplit(".", 1)[0] + "." if "." in name else ""
                key = prefix + t.shared_from
            if key not in state:
                missing.append(name)
                continue
            src = state[key]
            if not isinstance(src, torch.Tensor):
                src = torch.as_tensor(src)
            d = tp_dim_of(t)
            logical = list(t.shape)
            if d is not None:
                logical[d] *= topo.tensor_parallel_size
            if list(src.shape) != logical:
                mismatched.append((name, tuple(src.shape), tuple(logical)))
                continue
            t.copy_(shard_tp(src, d).to(device=t.device, dtype=t.dtype))
    if strict and (missing or unexpected or mismatched):
        raise RuntimeError(
            f"load_full_state_dict(strict=True): missing={missing} unexpected={unexpected} "
            f"mismatched={mismatched}"
        )
    return missing, unexpected, mismatched


def owned_parameters(module: torch.nn.Module) -> Iterable[torch.nn.Parameter]:
    for p in module.parameters():
        if is_materialized(p):
            yield p
