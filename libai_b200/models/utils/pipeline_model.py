"""Stage-wise execution protocol shared by all pipeline-capable models.

The reference tags module blocks with stage ids (``set_pipeline_stage_id`` → nn.Graph
``set_stage``; e.g. libai/models/gpt_model.py:359-401) and OneFlow's compiler cuts the graph.
Here every model exposes three pieces and the pipeline engine (``libai_b200/parallel/pipeline.py``)
moves the hidden state between stages with NCCL p2p:

* ``stage_pre(**batch)``      – embeddings etc.; runs on the stage owning layer 0
* ``stage_layers()``          – ordered blocks, each with a ``layer_idx``
* ``stage_post(hidden, **batch)`` – final norm / head / loss; runs on the stage owning layer −1

``forward_stage`` strings them together for the local stage; with pp == 1 it is the whole model.
Activation checkpointing (reference: graph_base.py:130-144, one checkpoint per TransformerLayer)
is applied around each block when ``self.activation_checkpoint`` is set.
"""
from __future__ import annotations

from typing import Any, Dict, Iterable

import torch
from torch.utils.checkpoint import checkpoint

from libai_b200.utils import distributed as dutil


class _GradBoundary(torch.autograd.Function):
    """Identity in forward.  Its backward runs once the backward of everything downstream of this point (block
    ``layer_idx`` and all later blocks) has been *issued*, i.e. the gradients of their parameters are final for this
    micro-batch — the optimizer uses the moment to start reducing that part of the flat gradient buffer over NVLink
    while the remaining backward keeps the tensor cores busy (``FlatOptimizer.on_grads_ready``)."""

    @staticmethod
    def forward(ctx, hidden, layer_idx, callback):
        ctx.layer_idx, ctx.callback = layer_idx, callback
        return hidden.view_as(hidden)

    @staticmethod
    def backward(ctx, grad):
        ctx.callback(ctx.layer_idx)
        return grad, None, None


class PipelineStageMixin:
    activation_checkpoint: bool = False
    # set by the trainer for the last micro-batch of a step (None = no early gradient reduction)
    grad_ready_callback = None
    grad_ready_layers = ()
    # ZeRO stage 2 / 3 (optim/zero_buckets.py): brackets every block — parameter all-gather before, gradient
    # reduce-scatter after its backward; set by DefaultTrainer.build_optimizer
    zero_hooks = None

    # ---- to be provided by the model -------------------------------------------------------
    def stage_pre(self, **batch):
        raise NotImplementedError

    def stage_layers(self) -> Iterable[torch.nn.Module]:
        raise NotImplementedError

    def stage_post(self, hidden, **batch):
        raise NotImplementedError

    def stage_layer_call(self, layer, hidden, batch: Dict[str, Any]):
        """How one block consumes the running hidden state (override for enc-dec models)."""
        return layer(hidden)

    # ---- generic driver -----------------------------------------------------------------------
    def forward_stage(self, batch: Dict[str, Any], hidden_in=None):
        topo = dutil.get_dist_util()
        if topo.pipeline_parallel_size > 1 and hidden_in is None and not torch.is_grad_enabled():
            from libai_b200.parallel import pipeline as _pl

            if not _pl.in_schedule():
                # plain ``model(**batch)`` under pipeline parallelism (evaluator, inference pipelines, ``test``): run
                # the whole pipelined forward and return the last stage's output on every rank
                return _pl.pipelined_forward(self, batch)
        hidden = self.stage_pre(**batch) if topo.is_first_stage else hidden_in
        use_ckpt = self.activation_checkpoint and torch.is_grad_enabled()
        cb = self.grad_ready_callback if torch.is_grad_enabled() else None
        for layer in self.stage_layers():
            if not topo.owns_layer(getattr(layer, "layer_idx", 0)):
                continue
            if cb is not None and torch.is_tensor(hidden) and getattr(layer, "layer_idx", -1) in self.grad_ready_layers:
                hidden = _GradBoundary.apply(hidden, layer.layer_idx, cb)
            zh = self.zero_hooks
            lidx = getattr(layer, "layer_idx", 0)
            if zh is not None:
                hidden = zh.wrap_before(hidden, lidx)
            if use_ckpt:
                hidden = checkpoint(self.stage_layer_call, layer, hidden, batch, use_reentrant=False)
            else:
                hidden = self.stage_layer_call(layer, hidden, batch)
            if zh is not None:
                hidden = zh.wrap_after(hidden, lidx)
        if topo.is_last_stage:
            return self.stage_post(hidden, **batch)
        return hidden
