from .bert_model import BertForClassification, BertForPreTraining, BertModel
from .build import build_graph, build_model
from .gpt_model import GPTForPreTraining, GPTModel
from .llama_model import LlamaForCausalLM, LlamaModel
from .resmlp import ResMLP
from .roberta_model import RobertaForCausalLM, RobertaForPreTraining, RobertaModel
from .swin_transformer import SwinTransformer
from .swin_transformer_v2 import SwinTransformerV2
from .t5_model import T5ForPreTraining, T5Model
from .vision_transformer import VisionTransformer

__all__ = [
    "build_model", "build_graph",
    "BertModel", "BertForPreTraining", "BertForClassification",
    "RobertaModel", "RobertaForPreTraining", "RobertaForCausalLM",
    "GPTModel", "GPTForPreTraining",
    "T5Model", "T5ForPreTraining",
    "VisionTransformer", "SwinTransformer", "SwinTransformerV2", "ResMLP",
    "LlamaModel", "LlamaForCausalLM",
]
