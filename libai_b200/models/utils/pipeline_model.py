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


class PipelineStageMixin:
    activation_checkpoint: bool = False

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
        hidden = self.stage_pre(**batch) if topo.is_first_stage else hidden_in
        use_ckpt = self.activation_checkpoint and torch.is_grad_enabled()
        for layer in self.stage_layers():
            if not topo.owns_layer(getattr(layer, "layer_idx", 0)):
                continue
            if use_ckpt:
                hidden = checkpoint(self.stage_layer_call, layer, hidden, batch, use_reentrant=False)
            else:
                hidden = self.stage_layer_call(layer, hidden, batch)
        if topo.is_last_stage:
            return self.stage_post(hidden, **batch)
        return hidden
