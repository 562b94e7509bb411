"""BERT-large pre-training on synthetic tokens: the reference's benchmark BERT (docs Benchmark.md:13-19: 24 layers,
hidden 1024, 16 heads, seq 512; BASELINE.json config "BERT-large pretrain TP=2 DP=4 bf16") with random-init weights."""
from libai_b200.config import LazyCall, OmegaConf
from libai_b200.data import build_nlp_test_loader, build_nlp_train_loader
from libai_b200.data.datasets import SyntheticBertDataset
from libai_b200.evaluation import PPLEvaluator

from .common.models.bert import pretrain_model as model
from .common.models.graph import graph
from .common.optim import optim
from .common.train import train

model.cfg.vocab_size = 30592          # 30522 padded to a multiple of 128
model.cfg.hidden_layers = 24
model.cfg.hidden_size = 1024
model.cfg.num_attention_heads = 16
model.cfg.intermediate_size = 4096
model.cfg.max_position_embeddings = 512

dataloader = OmegaConf.create()
dataloader.train = LazyCall(build_nlp_train_loader)(
    dataset=[LazyCall(SyntheticBertDataset)(vocab_size=30592, seq_length=512, num_samples=1 << 20, seed=1234)],
    num_workers=2,
)
dataloader.test = [
    LazyCall(build_nlp_test_loader)(
        dataset=LazyCall(SyntheticBertDataset)(vocab_size=30592, seq_length=512, num_samples=64, seed=4321),
        test_batch_size=4, num_workers=0,
    )
]

train.dist.pipeline_num_layers = model.cfg.hidden_layers
optim.lr = 1e-4
train.train_micro_batch_size = 16
train.test_micro_batch_size = 4
train.train_iter = 100
train.log_period = 10
train.amp.enabled = True
train.evaluation.enabled = False
train.evaluation.evaluator = LazyCall(PPLEvaluator)()
train.output_dir = "./output/bert_large_synthetic"
