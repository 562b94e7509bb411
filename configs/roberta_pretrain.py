"""RoBERTa pre-training recipe (reference configs/roberta_pretrain.py)."""
from libai_b200.config import LazyCall
from libai_b200.evaluation import PPLEvaluator

from .common.data.roberta_dataset import dataloader, tokenization
from .common.models.roberta import pretrain_model as model
from .common.models.graph import graph
from .common.optim import optim
from .common.train import train

vocab_file = "./data_test/roberta_data/roberta-vocab.json"
merge_files = "./data_test/roberta_data/roberta-merges.txt"
data_prefix = "./data_test/roberta_data/loss_compara_content_sentence"

tokenization.tokenizer.vocab_file = vocab_file
tokenization.tokenizer.merges_file = merge_files
for _ds in (dataloader.train.dataset[0], dataloader.test[0].dataset):
    _ds.data_prefix = data_prefix
    _ds.indexed_dataset.data_prefix = data_prefix

model.cfg.num_attention_heads = 12
model.cfg.hidden_size = 768
model.cfg.hidden_layers = 8

train.input_placement_device = "cpu"
train.dist.data_parallel_size = 8
train.dist.tensor_parallel_size = 1
train.dist.pipeline_parallel_size = 1
train.dist.pipeline_num_layers = model.cfg.hidden_layers
train.train_micro_batch_size = 2
train.amp.enabled = True
for _ds in dataloader.train.dataset:
    _ds.max_seq_length = model.cfg.max_position_embeddings
train.evaluation.evaluator = LazyCall(PPLEvaluator)()
train.output_dir = "output/roberta_output"
