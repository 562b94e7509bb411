"""``graph`` namespace, kept for config compatibility (reference configs/common/models/graph.py).
There is a single execution mode in libai_b200; ``enabled`` does not change semantics."""
from libai_b200.config import DictConfig, LazyCall
from libai_b200.models.utils import GraphBase

graph = dict(
    enabled=True,
    debug=-1,
    auto_parallel=dict(
        enabled=False,
        enable_auto_parallel_ignore_user_sbp_config=False,
        trunk_algo=True,
        sbp_collector=False,
    ),
    train_graph=LazyCall(GraphBase)(is_train=True),
    global_mode=dict(enabled=False),
    eval_graph=LazyCall(GraphBase)(is_train=False),
)
graph = DictConfig(graph)
