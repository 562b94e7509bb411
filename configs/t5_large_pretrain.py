"""T5 pre-training recipe (reference configs/t5_large_pretrain.py)."""
from libai_b200.config import LazyCall
from libai_b200.evaluation import PPLEvaluator

from .common.data.t5_dataset import dataloader, tokenization
from .common.models.t5 import pretrain_model as model
from .common.models.graph import graph
from .common.optim import optim
from .common.train import train

vocab_file = "./data_test/bert_data/bert-base-chinese-vocab.txt"
data_prefix = "./data_test/bert_data/loss_compara_content_sentence"

tokenization.tokenizer.vocab_file = vocab_file
for _ds in (dataloader.train.dataset[0], dataloader.test[0].dataset):
    _ds.data_prefix = data_prefix
    _ds.indexed_dataset.data_prefix = data_prefix

model.cfg.num_attention_heads = 12
model.cfg.hidden_size = 384
model.cfg.hidden_layers = 6

train.input_placement_device = "cpu"
train.dist.pipeline_num_layers = 2 * model.cfg.hidden_layers  # encoder + decoder
train.train_micro_batch_size = 16
train.amp.enabled = True

train.evaluation.evaluator = LazyCall(PPLEvaluator)()
train.output_dir = "./output/t5_output"
