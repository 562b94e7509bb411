"""``GraphBase`` compatibility wrapper.

In the reference this ``nn.Graph`` subclass is the switchboard for AMP + GradScaler, gradient
accumulation, activation checkpointing, ZeRO, pipeline stage ids, op fusion flags, NCCL stream
placement and auto-parallel (libai/models/utils/graph_base.py:28-152).  Those features are
first-class in this framework (``DefaultTrainer`` / ``FlatOptimizer`` / ``PipelineSchedule1F1B``),
so ``GraphBase`` only records the requested options, applies the activation-checkpoint flag, and
forwards calls to the model — configs that reference it keep working.
"""
import logging

from torch import nn

logger = logging.getLogger(__name__)


class GraphBase(nn.Module):
    def __init__(self, model, optimizer=None, lr_scheduler=None, fp16=False, activation_checkpoint=False,
                 grad_acc_steps=1, zero_optim=False, zero_stage=0, is_train=True, auto_parallel_conf=None,
                 global_mode=None, **kwargs):
        super().__init__()
        self.model = model
        self.is_train = is_train
        self.options = dict(fp16=fp16, activation_checkpoint=activation_checkpoint, grad_acc_steps=grad_acc_steps,
                            zero_optim=zero_optim, zero_stage=zero_stage)
        if auto_parallel_conf is not None and getattr(auto_parallel_conf, "enabled", False):
            logger.warning("graph.auto_parallel is a OneFlow compiler feature and is ignored by libai_b200")
        if activation_checkpoint:
            self.set_activation_checkpoint()

    def forward(self, **kwargs):
        return self.model(**kwargs)

    def build(self, **kwargs):
        return self.forward(**kwargs)

    def set_activation_checkpoint(self):
        setter = getattr(type(self.model), "set_activation_checkpoint", None)
        if setter is not None:
            setter(self.model)
        for m in self.model.modules():
            if hasattr(m, "activation_checkpoint"):
                m.activation_checkpoint = True

    def set_pipeline_stage_id(self):
        setter = getattr(type(self.model), "set_pipeline_stage_id", None)
        if setter is not None:
            setter(self.model)
