"""Llama-2 7B shaped pre-training on synthetic tokens (BASELINE.json config "Llama-2 7B projects/Llama TP=4 DP=2 +
ZeRO-2"): 32 layers, hidden 4096, 32 heads, SwiGLU 11008, seq 2048, random-init weights."""
from libai_b200.config import LazyCall, OmegaConf
from libai_b200.data import build_nlp_test_loader, build_nlp_train_loader
from libai_b200.data.datasets import SyntheticGPTDataset
from libai_b200.evaluation import PPLEvaluator

from configs.common.models.graph import graph
from configs.common.models.llama import model
from configs.common.optim import optim
from configs.common.train import train

SEQ = 2048
model.cfg.max_position_embeddings = SEQ

dataloader = OmegaConf.create()
dataloader.train = LazyCall(build_nlp_train_loader)(
    dataset=[LazyCall(SyntheticGPTDataset)(vocab_size=32000, seq_length=SEQ, num_samples=1 << 20, seed=1234)],
    num_workers=2,
)
dataloader.test = [
    LazyCall(build_nlp_test_loader)(
        dataset=LazyCall(SyntheticGPTDataset)(vocab_size=32000, seq_length=SEQ, num_samples=16, seed=4321),
        test_batch_size=2, num_workers=0,
    )
]

train.dist.pipeline_num_layers = model.cfg.hidden_layers
optim.lr = 1e-5
train.train_micro_batch_size = 2
train.test_micro_batch_size = 1
train.train_iter = 100
train.log_period = 10
train.amp.enabled = True
train.evaluation.enabled = False
train.evaluation.evaluator = LazyCall(PPLEvaluator)()
train.zero_optimization.enabled = True
train.zero_optimization.stage = 2
train.output_dir = "./output/llama7b_synthetic"
