"""MagicPrompt fine-tuning recipe (reference projects/MagicPrompt/configs/gpt2_training.py)."""
from configs.common.models.graph import graph
from configs.common.optim import optim
from configs.common.train import train
from libai_b200.config import LazyCall
from libai_b200.evaluation import PPLEvaluator
from projects.MagicPrompt.configs.gpt2_dataset import dataloader, tokenization
from projects.MagicPrompt.configs.gpt2_inference import pretrain_model as model

model.cfg.update(embedding_dropout_prob=0.1, attention_dropout_prob=0.1, output_dropout_prob=0.1, amp_enabled=True,
                 bias_gelu_fusion=True, bias_dropout_fusion=True)
optim.lr = 5.0e-05
train.update(
    dict(
        output_dir="projects/MagicPrompt/oneflow_magicprompt", train_micro_batch_size=4, test_micro_batch_size=4,
        train_epoch=33, train_iter=10000, log_period=50, amp=dict(enabled=True), warmup_ratio=0,
        checkpointer=dict(period=8000, max_to_keep=20),
        dist=dict(data_parallel_size=1, tensor_parallel_size=1, pipeline_parallel_size=1,
                  pipeline_num_layers=model.cfg.hidden_layers),
        evaluation=dict(enabled=True, evaluator=LazyCall(PPLEvaluator)(), eval_iter=250, eval_period=4000),
        rdma_enabled=False,
    )
)
