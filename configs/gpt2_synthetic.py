"""GPT-2 pre-training on synthetic tokens: the benchmark model of the reference
(docs Benchmark.md:20-26: 24 layers, hidden 1024, 16 heads, seq 1024) with random-init weights."""
from libai_b200.config import LazyCall
from libai_b200.evaluation import PPLEvaluator

from .common.data.gpt_synthetic import dataloader
from .common.models.gpt import pretrain_model as model
from .common.models.graph import graph
from .common.optim import optim
from .common.train import train

model.cfg.vocab_size = 50304  # 50257 padded to a multiple of 128
model.cfg.hidden_layers = 24
model.cfg.hidden_size = 1024
model.cfg.ffn_hidden_size = 4096
model.cfg.num_attention_heads = 16
model.cfg.max_seq_length = 1024

for ds in dataloader.train.dataset:
    ds.vocab_size = model.cfg.vocab_size
    ds.seq_length = model.cfg.max_seq_length
dataloader.test[0].dataset.vocab_size = model.cfg.vocab_size
dataloader.test[0].dataset.seq_length = model.cfg.max_seq_length

train.dist.pipeline_num_layers = model.cfg.hidden_layers
optim.lr = 1.5e-4
train.train_micro_batch_size = 6
train.test_micro_batch_size = 4
train.train_iter = 100
train.log_period = 10
train.amp.enabled = True
train.evaluation.enabled = False
train.evaluation.evaluator = LazyCall(PPLEvaluator)()
train.evaluation.eval_metric = "lm_loss_PPL"
train.evaluation.eval_mode = "min"
train.output_dir = "./output/gpt2_synthetic"
