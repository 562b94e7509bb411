from .build import build_graph, build_model
from .gpt_model import GPTForPreTraining, GPTModel

__all__ = ["build_model", "build_graph", "GPTModel", "GPTForPreTraining"]
