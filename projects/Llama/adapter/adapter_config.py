"""LLaMA-Adapter config (reference projects/Llama/adapter/adapter_config.py)."""
from libai_b200.config import DictConfig, LazyCall, OmegaConf
from projects.Llama.adapter.adapter_model import LlamaForCausalLM
from projects.Llama.configs.llama_config import cfg as _base
from projects.Llama.tokenizer import LlamaTokenizer

cfg = DictConfig(dict(_base))
cfg.max_position_embeddings = 2048
cfg.adapter_len = 10
cfg.adapter_layer = 30
cfg.pretrained_model_path = "meta-llama/Llama-2-7b-hf/"

model = LazyCall(LlamaForCausalLM)(cfg=cfg)
tokenization = OmegaConf.create()
tokenization.make_vocab_size_divisible_by = 1
tokenization.tokenizer = LazyCall(LlamaTokenizer)(pretrained_model_path="Llama-2-7b-hf/tokenizer.model")
