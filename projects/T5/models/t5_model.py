"""T5 model of this project (reference projects/T5/models/t5_model.py): the relative-position-bias T5 shared with
projects/MT5 (``model_type="t5"`` → ReLU MLP, tied LM head)."""
from projects.MT5.mt5_model import (  # noqa: F401
    MT5Embedding as T5Embedding,
    MT5ForPreTraining as T5ForPreTraining,
    MT5Loss as T5Loss,
    MT5Model as T5Model,
    T5Attention,
    T5MLP,
    TransformerLayer,
)
