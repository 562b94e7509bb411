"""Spec: reference libai/models/build.py:19-53."""
from libai_b200.config import instantiate, try_get_key


def build_model(cfg):
    """Instantiate ``cfg`` (a lazy model record, or an already-built module)."""
    if "_target_" in cfg:
        return instantiate(cfg)
    raise ValueError("cfg.model must be a LazyCall record (contain `_target_`)")


def build_graph(cfg, model, optimizer=None, lr_scheduler=None, is_train=False):
    """API parity with the reference's nn.Graph builder: the switches it threads into the graph
    (``train.amp / activation_checkpoint / zero_optimization / num_accumulation_steps``) are
    consumed by ``DefaultTrainer`` directly; the eager model is the executable."""
    from .utils.graph_base import GraphBase

    return GraphBase(
        model, optimizer, lr_scheduler, is_train=is_train,
        fp16=try_get_key(cfg, "train.amp.enabled", default=False),
        activation_checkpoint=try_get_key(cfg, "train.activation_checkpoint.enabled", default=False),
        zero_optim=try_get_key(cfg, "train.zero_optimization.enabled", default=False),
        zero_stage=try_get_key(cfg, "train.zero_optimization.stage", default=0),
        grad_acc_steps=try_get_key(cfg, "train.num_accumulation_steps", default=1),
    )
