"""Parameter construction: TP-sharded, PP-placed, layout-invariant initialisation.

The reference allocates every weight as a *global* tensor with an SBP and a placement
(libai/layers/linear.py:98-105) — the logical tensor is initialised once and each rank keeps its
shard, so results do not depend on the parallel layout.  Here the same property is obtained
explicitly: every parameter draws from its own ``torch.Generator`` seeded by
``(base_seed, creation_index)``; the *full* logical tensor is initialised and the local TP shard is
sliced out.  Parameters of layers owned by another pipeline stage are created on the ``meta``
device (no memory, no RNG dependence).
"""
from __future__ import annotations

import inspect
import os
from contextlib import contextmanager
from typing import Callable, Optional, Sequence

import torch
from torch import nn

from libai_b200.parallel.state import mark_tp
from libai_b200.utils import distributed as dutil

_STATE = {"seed": 1234, "counter": 0, "dtype": torch.float32, "device": None}


def set_init_seed(seed: int) -> None:
    _STATE["seed"] = int(seed)
    _STATE["counter"] = 0


def default_param_dtype() -> torch.dtype:
    return _STATE["dtype"]


@contextmanager
def param_defaults(dtype: Optional[torch.dtype] = None, device=None, seed: Optional[int] = None):
    """Scope the dtype/device used by layers constructed inside (used by ``build_model``)."""
    old = dict(_STATE)
    if dtype is not None:
        _STATE["dtype"] = dtype
    if device is not None:
        _STATE["device"] = torch.device(device)
    if seed is not None:
        set_init_seed(seed)
    try:
        yield
    finally:
        _STATE["dtype"], _STATE["device"] = old["dtype"], old["device"]


def param_device() -> torch.device:
    return _STATE["device"] if _STATE["device"] is not None else dutil.get_device()


def skip_init() -> bool:
    """``LIBAI_B200_SKIP_INIT=1`` (reference: ONEFLOW_LINEAR_EMBEDDING_SKIP_INIT) leaves weights
    uninitialised – used by inference pipelines that load a checkpoint right after."""
    return os.getenv("LIBAI_B200_SKIP_INIT", os.getenv("ONEFLOW_LINEAR_EMBEDDING_SKIP_INIT", "0")) in ("1", "true", "True")


def _call_init(init_fn: Callable, tensor: torch.Tensor, gen: torch.Generator) -> None:
    try:
        params = inspect.signature(init_fn).parameters
    except (TypeError, ValueError):
        params = {}
    if "generator" in params:
        init_fn(tensor, generator=gen)
        return
    # init function without a generator argument: run it under a forked, re-seeded global RNG
    devices = [tensor.device] if tensor.is_cuda else []
    with torch.random.fork_rng(devices=devices):
        torch.manual_seed(gen.initial_seed())
        init_fn(tensor)


def create_parameter(
    full_shape: Sequence[int],
    init_fn: Optional[Callable],
    *,
    tp_dim: Optional[int] = None,
    layer_idx: int = 0,
    dtype: Optional[torch.dtype] = None,
    requires_grad: bool = True,
    shared_with: Optional[torch.Tensor] = None,
) -> nn.Parameter:
    """Create the local shard of a logical parameter of shape ``full_shape``.

    ``shared_with``: another parameter whose initial value this one must replicate (tied weights
    living on two pipeline stages) – both draw from the same per-parameter seed."""
    topo = dutil.get_dist_util()
    if shared_with is not None:
        idx = shared_with.init_index
    else:
        idx = _STATE["counter"]
        _STATE["counter"] += 1
    dtype = dtype or _STATE["dtype"]
    tp = topo.tensor_parallel_size
    local_shape = list(full_shape)
    if tp_dim is not None and tp > 1:
        assert full_shape[tp_dim] % tp == 0, (
            f"dimension {tp_dim} of parameter shape {tuple(full_shape)} is not divisible by "
            f"tensor_parallel_size={tp}"
        )
        local_shape[tp_dim] //= tp
    if not topo.owns_layer(layer_idx):
        p = nn.Parameter(torch.empty(local_shape, dtype=dtype, device="meta"), requires_grad=requires_grad)
        p.init_index = idx
        return mark_tp(p, tp_dim)
    device = param_device()
    if init_fn is None or skip_init():
        data = torch.empty(local_shape, dtype=dtype, device=device)
        if init_fn is None:
            data.zero_()
    else:
        gen = torch.Generator(device=device)
        gen.manual_seed((_STATE["seed"] * 1000003 + idx * 7919) % (2 ** 63 - 1))
        full = torch.empty(tuple(full_shape), dtype=torch.float32, device=device)
        _call_init(init_fn, full, gen)
        if tp_dim is not None and tp > 1:
            n = local_shape[tp_dim]
            full = full.narrow(tp_dim, topo.tp_rank * n, n)
        data = full.to(dtype).contiguous()
    p = nn.Parameter(data, requires_grad=requires_grad)
    p.init_index = idx
    return mark_tp(p, tp_dim)


# ---- initialisers that accept a generator -------------------------------------------------
def xavier_normal_(t, gain: float = 1.0, generator=None):
    fan_out, fan_in = (t.shape[0], t.shape[1]) if t.dim() >= 2 else (t.shape[0], t.shape[0])
    if t.dim() > 2:
        rf = t[0][0].numel()
        fan_in, fan_out = t.shape[1] * rf, t.shape[0] * rf
    std = gain * (2.0 / float(fan_in + fan_out)) ** 0.5
    return t.normal_(0.0, std, generator=generator)


def ones_(t, generator=None):
    return t.fill_(1.0)


def zeros_(t, generator=None):
    return t.zero_()


def normal_init(std: float, mean: float = 0.0):
    def init_(t, generator=None):
        return t.normal_(mean, std, generator=generator)

    return init_


def trunc_normal_(t, mean=0.0, std=1.0, a=-2.0, b=2.0, generator=None):
    return nn.init.trunc_normal_(t, mean=mean, std=std, a=a, b=b, generator=generator)
