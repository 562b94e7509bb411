"""Aquila instruction tuning recipe (reference projects/Aquila/configs/aquila_sft.py)."""
import os

from configs.common.models.graph import graph
from configs.common.optim import optim
from configs.common.train import train
from libai_b200.config import LazyCall, OmegaConf
from libai_b200.data.build import build_nlp_test_loader, build_nlp_train_loader
from libai_b200.evaluation import PPLEvaluator
from libai_b200.scheduler import WarmupExponentialLR
from projects.Aquila.configs.aquila_config import cfg, tokenization
from projects.Aquila.aquila_dataset import AquilaDataset
from projects.Aquila.aquila import AquilaForCausalLM

dataset_path = "./data_aquila"
graph["enabled"] = False
optim.update(dict(lr=5e-5, weight_decay=0.1))

model = LazyCall(AquilaForCausalLM)(cfg=cfg)

dataloader = OmegaConf.create()
dataloader.train = LazyCall(build_nlp_train_loader)(
    dataset=[LazyCall(AquilaDataset)(path=os.path.join(dataset_path, "train"), tokenizer=tokenization.tokenizer)],
)
dataloader.test = [
    LazyCall(build_nlp_test_loader)(
        dataset=LazyCall(AquilaDataset)(path=os.path.join(dataset_path, "test"), tokenizer=tokenization.tokenizer),
    ),
]

train.update(
    dict(
        output_dir="./sft_result", train_micro_batch_size=2, test_micro_batch_size=1, train_epoch=3, train_iter=1,
        log_period=10, warmup_ratio=1 / 3, num_accumulation_steps=8, rdma_enabled=False, amp=dict(enabled=True),
        activation_checkpoint=dict(enabled=True), checkpointer=dict(period=5000, max_to_keep=20),
        dist=dict(data_parallel_size=1, tensor_parallel_size=1, pipeline_parallel_size=8,
                  pipeline_num_layers=cfg.hidden_layers),
        evaluation=dict(enabled=False, evaluator=LazyCall(PPLEvaluator)(), eval_period=1000, eval_iter=1e5),
        scheduler=LazyCall(WarmupExponentialLR)(warmup_factor=0.0, gamma=1.0, warmup_method="linear"),
    )
)
