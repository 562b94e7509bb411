"""Layer classes of the project (the reference keeps them in separate files under projects/MT5/layers/)."""
from projects.MT5.mt5_model import (  # noqa: F401
    MT5Embedding,
    MT5Loss,
    MT5MLP,
    T5Attention as MultiheadAttention,
    T5MLP,
    TransformerLayer,
    relative_position_bucket,
)
