"""ZeRO stage 2 and 3: gradients (stage 2) and parameters (stage 3) partitioned over the data-parallel group.

Reference switches: ``graph.config.enable_zero(True, stage)`` (libai/models/utils/graph_base.py:69-70; stage 3 is what
the reference's own tests run, tests/models/test_gpt.py:185-199) — OneFlow's compiler does the partitioning there.
Here it is explicit and bucketed **per transformer block**:

* every block's parameters form one *bucket* (padded to a multiple of the DP size); rank ``r`` owns slice ``r`` of every
  bucket: fp32 master weights, optimizer moments, the reduced-gradient shard and (stage 3) the low-precision parameter
  shard are the only per-bucket tensors that live for the whole run — ``1/dp`` of the bucket each;
* **stage 2**: the full fp32 gradient buffer of a block is live only while that block's backward runs: it is opened
  (zeroed) when the gradient reaches the block's output and reduce-scattered into the owners' shards as soon as the
  block's backward has been issued.  The storage comes from a **pool of two buffers per bucket shape** (even / odd
  blocks): a block's gradients therefore always land at the same addresses — which is what lets the blocks' backward
  be replayed from CUDA graphs (engine/cuda_graphs.py) — and the reduce-scatter of block ``i`` (asynchronous, on NCCL's
  stream) overlaps the backward of block ``i-1``, which writes into the other buffer.  Gradient memory is two block
  buckets instead of the whole model.  Full low-precision parameters stay resident (all-gathered after the optimizer
  step, like stage 1);
* **stage 3**: additionally the full parameters of a block exist only around its forward and its backward: all-gathered
  from the shards right before, dropped right after (the autograd nodes keep the ``Parameter`` objects, whose storage is
  simply re-pointed, so nothing stale is ever read);
* parameters outside the blocks (embeddings, final norm, heads — tied weights receive gradients at both ends of the
  backward pass) form *persistent* buckets: full parameters and full gradients stay resident, their gradients are
  reduce-scattered at the end of the backward pass.

The collectives are ``torch.distributed`` reduce-scatter / all-gather on the DP group (NCCL on GPUs, gloo on CPU), issued
from autograd hooks in ``PipelineStageMixin.forward_stage`` — outside the blocks, so stage 2 keeps the CUDA-graph replay
of the blocks; stage 3 does not (the parameter storage comes and goes).  Stage 1 keeps its fused NVLink kernels
(``optimizers.py``).
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.distributed as dist

from libai_b200.utils import distributed as dutil

_ALIGN = 256


class Bucket:
    """One parameter bucket; duck-types ``optimizers._Group`` for the update / norm / (de)serialisation code."""

    sharded = True
    symm = None

    def __init__(self, params: List[torch.nn.Parameter], dp_size: int, dp_rank: int, stage: int, persistent: bool, key,
                 grad_pool: Optional[Dict] = None, parity: int = 0):
        self.params, self.key, self.stage, self.persistent = params, key, stage, persistent
        self.pool_buf = None
        self.pool_part = None
        self.dtype, self.device = params[0].dtype, params[0].device
        assert all(p.dtype == self.dtype for p in params), "mixed dtypes inside one bucket"
        self.offsets, n = [], 0
        for p in params:
            self.offsets.append(n)
            n = (n + p.numel() + 7) // 8 * 8
        chunk = _ALIGN * dp_size
        self.numel = (n + chunk - 1) // chunk * chunk
        self.per = self.numel // dp_size
        self.lo, self.hi = dp_rank * self.per, (dp_rank + 1) * self.per
        full = torch.zeros(self.numel, dtype=self.dtype, device=self.device)
        for p, off in zip(params, self.offsets):
            full[off : off + p.numel()].copy_(p.data.reshape(-1))
        # persistent per-rank state: 1/dp of the bucket each (buckets whose full parameters stay resident use the owned
        # slice of the full buffer itself as the shard — no second copy)
        resident = stage < 3 or persistent
        self.param_shard = full[self.lo : self.hi] if resident else full[self.lo : self.hi].clone()
        self.master = self.param_shard.float() if self.dtype != torch.float32 else None
        self.red = torch.zeros(self.per, dtype=torch.float32, device=self.device)      # reduced-gradient shard
        self.state: Dict[str, torch.Tensor] = {}
        self._attrs = [{k: getattr(p, k) for k in ("tp_dim", "tp_stride", "sequence_parallel", "init_index", "shared_from",
                                                   "is_tied_source") if hasattr(p, k)} for p in params]
        self.param_flat: Optional[torch.Tensor] = None
        self.grad_flat: Optional[torch.Tensor] = None
        self.params_resident = stage < 3 or persistent
        self.grads_resident = persistent
        if not persistent and grad_pool is not None:
            # pooled gradient storage: buckets of this shape in even (odd) blocks share one buffer — every block's
            # gradients always land at the same addresses
            k = (self.numel, parity, str(self.device))
            if k not in grad_pool:
                grad_pool[k] = torch.zeros(self.numel, dtype=torch.float32, device=self.device)
                # the reduce-scatter output of the same parity: persistent too — a fresh `torch.empty` per block would be
                # handed to NCCL's stream (record_stream) and could not be reused by the caching allocator until that
                # stream caught up: with 7B-sized buckets (hundreds of MB) that means cudaMalloc / cudaFree in the step
                grad_pool[("part",) + k] = torch.empty(self.per, dtype=torch.float32, device=self.device)
            self.pool_buf = grad_pool[k]
            self.pool_part = grad_pool[("part",) + k]
        if self.params_resident:
            self._bind_params(full)
        else:
            self._unbind_params()
        if self.grads_resident:
            self._bind_grads()
        else:
            for p in params:
                p.main_grad = None
                p.grad_added_to_main_grad = False

    # ---- views -----------------------------------------------------------------------------------------------------
    def _rebind(self, p, i, data):
        p.data = data
        for k, v in self._attrs[i].items():
            setattr(p, k, v)

    def _bind_params(self, full: torch.Tensor):
        self.param_flat = full
        for i, (p, off) in enumerate(zip(self.params, self.offsets)):
            self._rebind(p, i, full[off : off + p.numel()].view(p.shape))

    def _unbind_params(self):
        self.param_flat = None
        for i, p in enumerate(self.params):
            # keep the logical shape visible (state_dict keys / shape checks) without holding memory
            self._rebind(p, i, torch.empty(1, dtype=self.dtype, device=self.device).expand(p.shape))

    def _bind_grads(self):
        if self.pool_buf is None:
            self.grad_flat = torch.zeros(self.numel, dtype=torch.float32, device=self.device)
        else:
            self.pool_buf.zero_()
            self.grad_flat = self.pool_buf
        for p, off in zip(self.params, self.offsets):
            p.main_grad = self.grad_flat[off : off + p.numel()].view(p.shape)
            p.grad_added_to_main_grad = False

    # ---- _Group interface ------------------------------------------------------------------------------------------
    def master_view(self) -> torch.Tensor:
        return self.master if self.master is not None else self.param_shard

    def grad_shard(self) -> torch.Tensor:
        return self.red

    def lp_shard(self) -> torch.Tensor:
        return self.param_shard

    def reduced_view(self, lo: int, hi: int) -> torch.Tensor:
        return self.red[lo - self.lo : hi - self.lo]

    # ---- parameter residency (stage 3) -------------------------------------------------------------------------------
    def gather_params(self, group):
        if self.param_flat is not None:
            return
        full = torch.empty(self.numel, dtype=self.dtype, device=self.device)
        _all_gather(full, self.param_shard, group)
        self._bind_params(full)

    def release_params(self):
        if not self.params_resident and self.param_flat is not None:
            self._unbind_params()

    def publish_params(self, group):
        """After the optimizer wrote ``param_shard``: refresh the resident full copy (if any)."""
        if self.param_flat is not None:
            _all_gather(self.param_flat, self.param_shard, group)

    def adopt_param_values(self):
        """Weights were written into the (materialised) full buffer behind the optimizer's back (checkpoint load):
        take this rank's slice as the new shard / master."""
        assert self.param_flat is not None
        self.param_shard.copy_(self.param_flat[self.lo : self.hi])
        if self.master is not None:
            self.master.copy_(self.param_shard)

    # ---- gradient residency (stage >= 2) -----------------------------------------------------------------------------
    def open_grads(self):
        if self.grad_flat is None:
            self._bind_grads()

    def reduce_grads(self, topo, scale: float, pending: Optional[list] = None):
        """Fold autograd gradients, apply the tensor-parallel fix-ups, reduce-scatter over DP into ``red``.  With
        ``pending`` (a list) the reduce-scatter is left in flight on NCCL's stream: ``(work, bucket, part)`` is appended
        and ``finish_reduce`` must be called before the bucket's pool buffer is opened again."""
        if self.grad_flat is None:
            return
        for p in self.params:
            if p.grad is not None:
                p.main_grad.add_(p.grad.to(torch.float32))
                p.grad = None
        if topo.sequence_parallel and topo.tp_group is not None:
            sp = [p.main_grad.view(-1) for p in self.params if getattr(p, "sequence_parallel", False)]
            if sp:
                packed = torch.cat(sp) if len(sp) > 1 else sp[0]
                dist.all_reduce(packed, group=topo.tp_group)
                if len(sp) > 1:
                    torch._foreach_copy_(sp, list(packed.split([g.numel() for g in sp])))
        if topo.embedding_group is not None:
            for p in self.params:
                if getattr(p, "shared_from", None) is not None or getattr(p, "is_tied_source", False):
                    dist.all_reduce(p.main_grad, group=topo.embedding_group)
        if scale != 1.0:
            self.grad_flat.mul_(scale)
        part = self.pool_part if self.pool_part is not None else torch.empty(self.per, dtype=torch.float32, device=self.device)
        if pending is not None and self.grad_flat.is_cuda and not self.grads_resident:
            work = dist.reduce_scatter_tensor(part, self.grad_flat, group=topo.dp_group, async_op=True)
            pending.append((work, self, part))
        else:
            _reduce_scatter(part, self.grad_flat, topo.dp_group, topo.dp_rank)
            self.red.add_(part)
        if self.grads_resident:
            self.grad_flat.zero_()
        else:
            # closed: eager kernels fall back to autograd's ``p.grad`` until the bucket is opened again (captured
            # kernels keep the pool address they were recorded with)
            self.grad_flat = None
            for p in self.params:
                p.main_grad = None
                p.grad_added_to_main_grad = False

    @staticmethod
    def finish_reduce(pending: list):
        while pending:
            work, bucket, part = pending.pop(0)
            work.wait()                      # stream-level wait: the current stream continues after the collective
            bucket.red.add_(part)


def _all_gather(full: torch.Tensor, shard: torch.Tensor, group):
    if full.is_cuda:
        dist.all_gather_into_tensor(full, shard.contiguous(), group=group)
    else:   # gloo
        world = dist.get_world_size(group)
        parts = [torch.empty_like(shard) for _ in range(world)]
        dist.all_gather(parts, shard.contiguous(), group=group)
        full.copy_(torch.cat(parts))


def _reduce_scatter(out: torch.Tensor, full: torch.Tensor, group, rank: int):
    if full.is_cuda:
        dist.reduce_scatter_tensor(out, full, group=group)
    else:   # gloo has no reduce-scatter
        tmp = full.clone()
        dist.all_reduce(tmp, group=group)
        n = out.numel()
        out.copy_(tmp[rank * n : (rank + 1) * n])


class _Before(torch.autograd.Function):
    """Forward: make the block's parameters resident.  Backward (runs once the block's backward has been issued): its
    gradient bucket is complete → reduce-scatter it, drop the bucket and (stage 3) the parameters."""

    @staticmethod
    def forward(ctx, hidden, hooks, idx):
        ctx.hooks, ctx.idx = hooks, idx
        hooks.pre_forward(idx)
        return hidden.view_as(hidden)

    @staticmethod
    def backward(ctx, grad):
        ctx.hooks.post_backward(ctx.idx)
        return grad, None, None


class _After(torch.autograd.Function):
    """Forward: the block is done, (stage 3) drop its parameters.  Backward (runs right before the block's backward):
    open the gradient bucket and (stage 3) gather the parameters again."""

    @staticmethod
    def forward(ctx, hidden, hooks, idx):
        ctx.hooks, ctx.idx = hooks, idx
        hooks.post_forward(idx)
        return hidden.view_as(hidden)

    @staticmethod
    def backward(ctx, grad):
        ctx.hooks.pre_backward(ctx.idx)
        return grad, None, None


class ZeroBucketHooks:
    """What ``forward_stage`` calls around every block (``wrap``), see ``_Before`` / ``_After``."""

    def __init__(self, by_layer: Dict[int, List[Bucket]], dp_grad_reduce: str, stage: int = 2):
        self.by_layer = by_layer
        self.scale = 1.0
        self.dp_grad_reduce = dp_grad_reduce
        self.stage = stage
        self._pending: list = []       # reduce-scatters in flight (at most one block's)

    @property
    def graph_capturable(self) -> bool:
        """Stage 2: parameters resident, pooled gradient buffers at fixed addresses → the blocks can be replayed from
        CUDA graphs.  Stage 3 re-points the parameter storage around every block."""
        return self.stage == 2

    def open_all(self):
        """Bind every block's gradient views (capture / warm-up runs the blocks' backward outside ``forward_stage``)."""
        for buckets in self.by_layer.values():
            for b in buckets:
                b.open_grads()

    def close_all(self):
        for buckets in self.by_layer.values():
            for b in buckets:
                if not b.grads_resident and b.grad_flat is not None:
                    b.grad_flat = None
                    for p in b.params:
                        p.main_grad = None
                        p.grad_added_to_main_grad = False

    def flush(self):
        """Wait for the reduce-scatter still in flight (end of the backward pass)."""
        Bucket.finish_reduce(self._pending)

    def _topo(self):
        return dutil.get_dist_util()

    def wrap_before(self, hidden, idx):
        if idx in self.by_layer and torch.is_tensor(hidden):
            if hidden.requires_grad and torch.is_grad_enabled():
                return _Before.apply(hidden, self, idx)
            self.pre_forward(idx)
        return hidden

    def wrap_after(self, hidden, idx):
        if idx in self.by_layer and torch.is_tensor(hidden):
            if hidden.requires_grad and torch.is_grad_enabled():
                return _After.apply(hidden, self, idx)
            self.post_forward(idx)
        return hidden

    def pre_forward(self, idx):
        topo = self._topo()
        for b in self.by_layer[idx]:
            b.gather_params(topo.dp_group)

    def post_forward(self, idx):
        for b in self.by_layer[idx]:
            b.release_params()

    def pre_backward(self, idx):
        topo = self._topo()
        for b in self.by_layer[idx]:
            b.gather_params(topo.dp_group)
            b.open_grads()

    def post_backward(self, idx):
        topo = self._topo()
        scale = 1.0 / topo.data_parallel_size if self.dp_grad_reduce == "mean" else 1.0
        # the previous block's reduce-scatter ran under this block's backward; it must be done before ITS pool buffer is
        # opened again (two blocks further down) — finishing it here keeps at most one in flight
        self.flush()
        for b in self.by_layer[idx]:
            b.reduce_grads(topo, scale, pending=self._pending)
            b.release_params()


def build_buckets(param_groups, model, dp_size: int, dp_rank: int, stage: int):
    """Split every optimizer param group into buckets: one per transformer block that owns parameters of the group, plus
    one persistent bucket for everything else.  Returns ``(buckets_per_group, by_layer)``."""
    layer_of = {}
    layers = model.stage_layers() if (model is not None and hasattr(model, "stage_layers")) else []
    try:
        layers = list(layers)
    except TypeError:
        layers = []
    for k, layer in enumerate(layers):
        idx = getattr(layer, "layer_idx", k)
        for p in layer.parameters():
            layer_of.setdefault(id(p), idx)
    per_group, by_layer = [], {}
    order = {idx: k for k, idx in enumerate(sorted({v for v in layer_of.values()}))}
    for gi, g in enumerate(param_groups):
        pool: Dict = {}
        params = [p for p in g["params"] if p.device.type != "meta" and p.requires_grad]
        split: Dict[object, List] = {}
        for p in params:
            split.setdefault(layer_of.get(id(p), "rest"), []).append(p)
        buckets = []
        for key in sorted(split, key=lambda k: (isinstance(k, str), k)):
            b = Bucket(split[key], dp_size, dp_rank, stage, persistent=(key == "rest"), key=key,
                       grad_pool=pool, parity=order.get(key, 0) % 2)
            buckets.append(b)
            if key != "rest":
                by_layer.setdefault(key, []).append(b)
        per_group.append(buckets)
    return per_group, by_layer
