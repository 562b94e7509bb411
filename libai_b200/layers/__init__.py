from .activation import build_activation
from .attention import AttnMaskType, MultiheadAttention
from .conv import Conv1D
from .cross_entropy import ParallelCrossEntropyLoss
from .droppath import DropPath
from .embedding import Embedding, PatchEmbedding, SinePositionalEmbedding, VocabEmbedding
from .layer_norm import LayerNorm, RMSLayerNorm
from .linear import Linear, Linear1D
from .lm_logits import LMLogits
from .mlp import MLP
from .transformer_layer import TransformerLayer

__all__ = [
    "Embedding", "VocabEmbedding", "SinePositionalEmbedding", "PatchEmbedding", "build_activation",
    "Linear", "Linear1D", "Conv1D", "MLP", "LayerNorm", "RMSLayerNorm", "TransformerLayer",
    "MultiheadAttention", "AttnMaskType", "ParallelCrossEntropyLoss", "LMLogits", "DropPath",
]
