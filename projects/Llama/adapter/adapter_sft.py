"""Adapter tuning recipe (reference projects/Llama/adapter/adapter_sft.py): the SFT recipe with the adapter model,
a higher learning rate and no weight decay on the (few) trainable tensors."""
from projects.Llama.adapter.adapter_config import cfg
from projects.Llama.adapter.adapter_model import LlamaForCausalLM
from projects.Llama.configs.llama_sft import dataloader, graph, optim, tokenization, train
from libai_b200.config import LazyCall

model = LazyCall(LlamaForCausalLM)(cfg=cfg)
optim.update(dict(lr=9e-3, weight_decay=0.02))
train.update(dict(output_dir="./adapter_result", train_epoch=5, warmup_ratio=2 / 5,
                  dist=dict(data_parallel_size=1, tensor_parallel_size=1, pipeline_parallel_size=8,
                            pipeline_num_layers=cfg.hidden_layers)))
