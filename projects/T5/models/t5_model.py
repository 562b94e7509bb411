"""T5 model of this project (reference projects/T5/models/t5_model.py): the relative-position-bias T5 shared with
projects/MT5 (``model_type="t5"`` → ReLU MLP, tied LM head)."""
from projects.MT5.mt5_model import (  # noqa: F401
    MT5Embedding as T5Embedding,
    MT5ForPreTraining as T5ForPreTraining,
    MT5Loss as T5Loss,
    MT5Model as T5Model,
    MT5MLP,
    T5Attention,
    T5MLP,
    TransformerLayer,
)
from libai_b200.layers import Embedding, LMLogits  # noqa: F401
from libai_b200.layers import RMSLayerNorm as LayerNorm  # noqa: F401  (T5 normalises without mean / bias)

MultiheadAttention = T5Attention
