from .base_loader import ModelLoaderHuggerFace, ModelLoaderLiBai
from .bert_loader import BertLoaderHuggerFace, BertLoaderLiBai
from .gpt_loader import GPT2LoaderHuggerFace, GPT2LoaderLiBai
from .llama_loader import LlamaLoaderHuggerFace, LlamaLoaderLiBai
from .roberta_loader import RobertaLoaderHuggerFace, RobertaLoaderLiBai
from .swin_loader import SwinLoaderHuggerFace, SwinLoaderLiBai
from .swinv2_loader import SwinV2LoaderHuggerFace, SwinV2LoaderLiBai
from .vit_loader import ViTLoaderHuggerFace, ViTLoaderLiBai
