"""ChatGLM SFT / LoRA recipe (reference projects/ChatGLM/configs/chatglm_sft.py)."""
import os

from configs.common.models.graph import graph
from configs.common.optim import optim
from configs.common.train import train
from libai_b200.config import LazyCall, OmegaConf
from libai_b200.data.build import build_nlp_test_loader, build_nlp_train_loader
from libai_b200.evaluation import PPLEvaluator
from libai_b200.scheduler import WarmupExponentialLR
from projects.ChatGLM.chatglm import ChatGLMForConditionalGeneration
from projects.ChatGLM.configs.chatglm_config import cfg, tokenization
from projects.ChatGLM.dataset import ChatGLMTrainDataset

lora_enable = os.getenv("CHATGLM_LORA", "0") == "1"
dataset_path = os.getenv("DATA_DIR", "./data/alpaca")
max_source_len, max_target_len = 128, 128

graph["enabled"] = False
optim.update(dict(lr=2e-5 if not lora_enable else 1e-4, weight_decay=0.1))
cfg.lora_enable = lora_enable

model = LazyCall(ChatGLMForConditionalGeneration)(cfg=cfg)

dataloader = OmegaConf.create()
dataloader.train = LazyCall(build_nlp_train_loader)(
    dataset=[LazyCall(ChatGLMTrainDataset)(path=os.path.join(dataset_path, "train.json"), tokenizer=tokenization.tokenizer,
                                            max_source_len=max_source_len, max_target_len=max_target_len)],
)
dataloader.test = [
    LazyCall(build_nlp_test_loader)(
        dataset=LazyCall(ChatGLMTrainDataset)(path=os.path.join(dataset_path, "test.json"), tokenizer=tokenization.tokenizer,
                                              max_source_len=max_source_len, max_target_len=max_target_len),
    ),
]

train.update(
    dict(
        output_dir="./sft_result", train_micro_batch_size=1, test_micro_batch_size=1, train_epoch=3, train_iter=1,
        log_period=10, warmup_ratio=2 / 5, num_accumulation_steps=8, rdma_enabled=True, amp=dict(enabled=True),
        activation_checkpoint=dict(enabled=True), checkpointer=dict(period=5000, max_to_keep=1),
        dist=dict(data_parallel_size=1, tensor_parallel_size=1, pipeline_parallel_size=4, pipeline_num_layers=cfg.num_layers),
        evaluation=dict(enabled=False, evaluator=LazyCall(PPLEvaluator)(), eval_period=1000, eval_iter=1e5),
        scheduler=LazyCall(WarmupExponentialLR)(warmup_factor=0.0, gamma=1.0, warmup_method="linear"),
    )
)
