"""Pipeline parallelism: non-interleaved 1F1B schedule in user space.

The reference gets 1F1B from OneFlow's graph compiler (``set_stage`` per block,
``set_gradient_accumulation_steps`` = number of micro-batches; libai/models/gpt_model.py:359-401,
libai/models/utils/graph_base.py:63-64, docs customize_parallel.md:171-185).  Here the schedule is
explicit:

    stage s of p, M micro-batches:   warm-up  = min(p - s - 1, M) forwards
                                      steady   = M - warm-up  × (forward, backward)
                                      cooldown = warm-up backwards

Activations / activation-gradients travel between neighbouring stages (same dp, tp coordinates) with
NCCL ``isend/irecv`` (gloo on CPU).  Under sequence parallelism the payload is the token shard
``[b·s/t, h]`` – t× smaller than the reference's replicated ``[b, s, h]`` hand-off.  Models implement
``forward_stage(batch, hidden_in)`` (see models/utils/pipeline_model.py).  Tied embeddings: the
first and last stage hold separate copies whose gradients are summed over ``embedding_group`` by
the optimizer's ``sync_gradients``.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple, Union

import torch
import torch.distributed as dist

from libai_b200.layers.embedding import set_sp_shape
from libai_b200.utils import distributed as dutil

_DTYPES = [torch.float32, torch.bfloat16, torch.float16, torch.int64, torch.int32]
TensorOrTuple = Union[torch.Tensor, Tuple[torch.Tensor, ...]]


def _as_tuple(x) -> Tuple[torch.Tensor, ...]:
    return tuple(x) if isinstance(x, (tuple, list)) else (x,)


class _P2P:
    """Shape-aware send/recv between pipeline neighbours (shape/dtype handshake on first use)."""

    def __init__(self):
        self.topo = dutil.get_dist_util()
        self.group = self.topo.pp_group
        self.prev = self.topo.pp_ranks[self.topo.pp_rank - 1] if not self.topo.is_first_stage else None
        self.next = self.topo.pp_ranks[self.topo.pp_rank + 1] if not self.topo.is_last_stage else None
        self.meta_fwd: Optional[List[Tuple[Tuple[int, ...], torch.dtype]]] = None  # what we receive from prev
        self.sent_meta = False
        self.dev = self.topo.device

    # -- meta ------------------------------------------------------------------------------------
    def _send_meta(self, tensors: Sequence[torch.Tensor]):
        buf = torch.full((1 + 10 * len(tensors),), -1, dtype=torch.int64)
        buf[0] = len(tensors)
        for i, t in enumerate(tensors):
            buf[1 + i * 10] = _DTYPES.index(t.dtype)
            buf[2 + i * 10] = t.dim()
            for d, n in enumerate(t.shape):
                buf[3 + i * 10 + d] = n
        head = torch.tensor([buf.numel()], dtype=torch.int64, device=self.dev)
        dist.send(head, self.next, group=self.group)
        dist.send(buf.to(self.dev), self.next, group=self.group)

    def _recv_meta(self):
        head = torch.empty(1, dtype=torch.int64, device=self.dev)
        dist.recv(head, self.prev, group=self.group)
        buf = torch.empty(int(head.item()), dtype=torch.int64, device=self.dev)
        dist.recv(buf, self.prev, group=self.group)
        buf = buf.cpu()
        metas = []
        for i in range(int(buf[0])):
            dt = _DTYPES[int(buf[1 + i * 10])]
            nd = int(buf[2 + i * 10])
            metas.append((tuple(int(x) for x in buf[3 + i * 10 : 3 + i * 10 + nd]), dt))
        self.meta_fwd = metas

    # -- low level ----------------------------------------------------------------------------------
    def _exchange(self, sends, recvs):
        """``sends``/``recvs``: lists of (tensor, peer). Issued as one batch so that opposite
        directions between two neighbours progress concurrently (no send/send deadlock)."""
        ops = [dist.P2POp(dist.isend, t, peer, self.group) for t, peer in sends]
        ops += [dist.P2POp(dist.irecv, t, peer, self.group) for t, peer in recvs]
        if not ops:
            return
        for req in dist.batch_isend_irecv(ops):
            req.wait()

    def _fwd_buffers(self):
        if self.meta_fwd is None:
            self._recv_meta()
        return [torch.empty(shape, dtype=dt, device=self.dev) for shape, dt in self.meta_fwd]

    @staticmethod
    def _wrap_inputs(bufs):
        for t in bufs:
            if t.dtype.is_floating_point:
                t.requires_grad_(True)
        return bufs[0] if len(bufs) == 1 else tuple(bufs)

    def _fwd_payload(self, out):
        ts = [t.detach().contiguous() for t in _as_tuple(out)]
        if not self.sent_meta:
            self._send_meta(ts)
            self.sent_meta = True
        return ts

    @staticmethod
    def _grad_payload(hidden_in):
        return [
            (t.grad if t.grad is not None else torch.zeros_like(t)).contiguous()
            for t in _as_tuple(hidden_in)
            if t.dtype.is_floating_point
        ]

    @staticmethod
    def _grad_buffers(out):
        return [torch.empty_like(t) if t.dtype.is_floating_point else None for t in _as_tuple(out)]

    # -- single direction ----------------------------------------------------------------------------
    def recv_forward(self):
        if self.prev is None:
            return None
        bufs = self._fwd_buffers()
        self._exchange([], [(b, self.prev) for b in bufs])
        return self._wrap_inputs(bufs)

    def send_forward(self, out):
        if self.next is None:
            return
        self._exchange([(t, self.next) for t in self._fwd_payload(out)], [])

    def recv_backward(self, out):
        if self.next is None:
            return None
        bufs = self._grad_buffers(out)
        self._exchange([], [(b, self.next) for b in bufs if b is not None])
        return tuple(bufs)

    def send_backward(self, hidden_in):
        if self.prev is None or hidden_in is None:
            return
        self._exchange([(g, self.prev) for g in self._grad_payload(hidden_in)], [])

    # -- fused directions (steady state of 1F1B) -----------------------------------------------------------
    def send_forward_recv_backward(self, out):
        if self.next is None:
            return None
        bufs = self._grad_buffers(out)
        self._exchange([(t, self.next) for t in self._fwd_payload(out)], [(b, self.next) for b in bufs if b is not None])
        return tuple(bufs)

    def send_backward_recv_forward(self, hidden_in):
        if self.prev is None:
            return None
        bufs = self._fwd_buffers()
        self._exchange([(g, self.prev) for g in self._grad_payload(hidden_in)], [(b, self.prev) for b in bufs])
        return self._wrap_inputs(bufs)


# ``forward_stage`` called while a schedule drives the stages = one hop of that schedule; called from anywhere else
# under pp > 1 (evaluator, inference pipelines, ``DefaultTrainer.test``: plain ``model(**batch)``) = a request for the
# whole pipelined forward, see ``pipelined_forward``.
_SCHEDULE_DEPTH = 0


def in_schedule() -> bool:
    return _SCHEDULE_DEPTH > 0


class _ScheduleScope:
    def __enter__(self):
        global _SCHEDULE_DEPTH
        _SCHEDULE_DEPTH += 1

    def __exit__(self, *exc):
        global _SCHEDULE_DEPTH
        _SCHEDULE_DEPTH -= 1
        return False


def broadcast_from_stage(out, stage: int, topo=None):
    """Hand stage ``stage``'s value (tensor / tuple / list / dict of tensors, ``None`` and python scalars allowed) to
    every stage of this rank's pipeline group.  Used by inference code that walks the layers stage by stage
    (``projects/MT5`` generation with a KV cache) — correctness first, one broadcast per stage boundary."""
    topo = topo or dutil.get_dist_util()
    if topo.pipeline_parallel_size == 1:
        return out
    return _broadcast_from_last_stage(out, topo, stage)


def _broadcast_from_last_stage(out, topo, stage: Optional[int] = None):
    """Hand the last stage's result (tensor / tuple / dict of tensors, nested python scalars allowed) to every stage
    of the pipeline group so that callers see the same value on all ranks."""
    stage = topo.pipeline_parallel_size - 1 if stage is None else stage
    src = topo.pp_ranks[stage]
    group = topo.pp_group
    dev = topo.device
    is_src = topo.pp_rank == stage

    def describe(o):
        if torch.is_tensor(o):
            return ("t", tuple(o.shape), o.dtype)
        if isinstance(o, dict):
            return ("d", [(k, describe(v)) for k, v in o.items()])
        if isinstance(o, (tuple, list)):
            return ("l" if isinstance(o, list) else "u", [describe(v) for v in o])
        return ("o", o)

    meta = [describe(out) if is_src else None]
    dist.broadcast_object_list(meta, src=src, group=group, device=dev if dev.type == "cuda" else None)

    def rebuild(m, o):
        kind = m[0]
        if kind == "t":
            t = o.detach().contiguous() if is_src else torch.empty(m[1], dtype=m[2], device=dev)
            if t.numel():
                dist.broadcast(t, src=src, group=group)
            return t
        if kind == "d":
            return {k: rebuild(mv, o[k] if is_src else None) for k, mv in m[1]}
        if kind in ("l", "u"):
            vals = [rebuild(mv, o[i] if is_src else None) for i, mv in enumerate(m[1])]
            return vals if kind == "l" else tuple(vals)
        return m[1]

    return rebuild(meta[0], out)


_INFER_SCHEDULES: Dict[int, "PipelineSchedule1F1B"] = {}


@torch.no_grad()
def pipelined_forward(model, batch: Dict[str, torch.Tensor]):
    """Whole-pipeline forward for evaluation / inference: every stage runs its blocks, activations hop stage to stage
    over NCCL/gloo p2p, and the last stage's output is broadcast back so all ranks return it (reference semantics: a
    global tensor readable everywhere, libai/evaluation/evaluator.py:130-190 ``to_global`` + ``to_local``)."""
    sched = _INFER_SCHEDULES.get(id(model))
    if sched is None:
        sched = _INFER_SCHEDULES[id(model)] = PipelineSchedule1F1B(model)
        sched.p2p = _P2PDynamic()     # shapes change from call to call (generation, last partial batch)
    out = sched.forward_only(batch)
    return _broadcast_from_last_stage(out, sched.topo)


class _P2PDynamic(_P2P):
    """Inference flavour: the shape/dtype handshake is repeated on every hop (payload shapes are not static)."""

    def _fwd_buffers(self):
        self._recv_meta()
        return [torch.empty(shape, dtype=dt, device=self.dev) for shape, dt in self.meta_fwd]

    def _fwd_payload(self, out):
        ts = [t.detach().contiguous() for t in _as_tuple(out)]
        self._send_meta(ts)
        return ts


class PipelineSchedule1F1B:
    """Runs one optimizer step's worth of micro-batches through the local pipeline stage."""

    def __init__(self, model, loss_scaler=None):
        self.model = model
        self.topo = dutil.get_dist_util()
        self.p2p = _P2P()
        self.loss_scaler = loss_scaler

    def _forward(self, batch: Dict[str, torch.Tensor], hidden_in, n_micro: int):
        """Returns ``(out, metrics)``: ``out`` is the scaled scalar loss on the last stage, the stage
        output tensors otherwise."""
        if self.topo.sequence_parallel:
            first = next(iter(batch.values()))
            set_sp_shape(first.shape[0], first.shape[1] if first.dim() > 1 else 1)
        with _ScheduleScope():
            out = self.model.forward_stage(batch, hidden_in)
        if self.topo.is_last_stage:
            loss = sum(v for k, v in out.items() if "loss" in k) / n_micro
            if self.loss_scaler is not None:
                loss = self.loss_scaler.scale_loss(loss)
            return loss, {k: v.detach() / n_micro for k, v in out.items()}
        return out, None

    def _backward(self, out, out_grads):
        if self.topo.is_last_stage:
            out.backward()
            return
        pairs = [(t, g) for t, g in zip(_as_tuple(out), out_grads) if g is not None and t.requires_grad]
        torch.autograd.backward([t for t, _ in pairs], [g for _, g in pairs])

    def run(self, batches: List[Dict[str, torch.Tensor]]) -> Optional[Dict[str, torch.Tensor]]:
        """Returns the (micro-batch averaged) loss dict on the last stage, ``None`` elsewhere."""
        M = len(batches)
        p, s = self.topo.pipeline_parallel_size, self.topo.pp_rank
        warm = min(p - s - 1, M)
        steady = M - warm
        p2p = self.p2p
        inflight: List[Tuple] = []  # (hidden_in, out) awaiting backward, FIFO
        metrics: Optional[Dict[str, torch.Tensor]] = None

        def track(m):
            nonlocal metrics
            if m is not None:
                metrics = m if metrics is None else {k: metrics[k] + m[k] for k in m}

        it = iter(batches)
        fwd_index = [0]

        def next_batch():
            # micro-batch i runs on captured copy i % n_slots of every block (engine/cuda_graphs.py)
            from libai_b200.engine.cuda_graphs import set_micro_batch_slot

            set_micro_batch_slot(fwd_index[0])
            fwd_index[0] += 1
            return next(it)

        # ---- warm-up forwards
        for _ in range(warm):
            hidden_in = p2p.recv_forward()
            out, m = self._forward(next_batch(), hidden_in, M)
            track(m)
            p2p.send_forward(out)
            inflight.append((hidden_in, out))
        # ---- steady 1F1B
        hidden_in = p2p.recv_forward() if steady > 0 else None
        for i in range(steady):
            out, m = self._forward(next_batch(), hidden_in, M)
            track(m)
            inflight.append((hidden_in, out))
            grads = p2p.send_forward_recv_backward(out)
            h_old, o_old = inflight.pop(0)
            self._backward(o_old, grads)
            if i == steady - 1:
                p2p.send_backward(h_old)
            else:
                hidden_in = p2p.send_backward_recv_forward(h_old)
                if self.topo.is_first_stage:
                    hidden_in = None
        # ---- cool-down backwards
        for _ in range(warm):
            h_old, o_old = inflight.pop(0)
            grads = p2p.recv_backward(o_old)
            self._backward(o_old, grads)
            p2p.send_backward(h_old)
        return metrics

    @torch.no_grad()
    def forward_only(self, batch: Dict[str, torch.Tensor]):
        """Inference / evaluation through the pipeline; the result lives on the last stage."""
        if self.topo.sequence_parallel:
            first = next((v for v in batch.values() if torch.is_tensor(v)), None)
            if first is not None:
                set_sp_shape(first.shape[0], first.shape[1] if first.dim() > 1 else 1)
        hidden_in = self.p2p.recv_forward()
        with _ScheduleScope():
            out = self.model.forward_stage(batch, hidden_in)
        if not self.topo.is_last_stage:
            self.p2p.send_forward(out)
            return None
        return out
