"""mT5 pre-training on the span-corruption T5 dataset: DP2 × TP2, ZeRO-2, 8 accumulation steps
(reference projects/MT5/configs/mt5_pretrain.py)."""
from configs.common.data.t5_dataset import dataloader, tokenization
from configs.common.models.graph import graph
from configs.common.optim import optim
from configs.common.train import train
from libai_b200.config import LazyCall
from libai_b200.evaluation import PPLEvaluator
from libai_b200.scheduler import WarmupExponentialLR
from projects.MT5.configs.mt5_base import pretrain_model as model

vocab_file = "./data_test/bert_data/bert-base-chinese-vocab.txt"
data_prefix = "./data_test/bert_data/loss_compara_content_sentence"

tokenization.tokenizer.vocab_file = vocab_file
dataloader.train.dataset[0].data_prefix = data_prefix
dataloader.train.dataset[0].indexed_dataset.data_prefix = data_prefix

model.cfg.update(
    hidden_size=768, hidden_layers=12, num_attention_heads=12, head_size=64, intermediate_size=2048, model_type="mt5",
    hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, embedding_dropout_prob=0.0, vocab_size=30522,
    padding_idx=0, tie_word_embeddings=False, is_encoder_decoder=False, amp_enabled=True, initializer_range=0.02,
    pretrained_model_path=None,
)

train.update(
    dict(
        output_dir="projects/MT5/output/mt5_output",
        train_micro_batch_size=4,
        train_epoch=1,
        train_iter=24000,
        log_period=10,
        amp=dict(enabled=True),
        warmup_ratio=1 / 24,
        input_placement_device="cpu",
        dist=dict(data_parallel_size=2, tensor_parallel_size=2, pipeline_parallel_size=1,
                  pipeline_num_layers=2 * model.cfg.hidden_layers),
        scheduler=LazyCall(WarmupExponentialLR)(warmup_factor=0.001, gamma=1.0, warmup_method="linear", warmup_iter=0.0),
        evaluation=dict(evaluator=LazyCall(PPLEvaluator)(), enabled=True, eval_iter=1e5, eval_period=5000),
    )
)
train.zero_optimization.enabled = True
train.zero_optimization.stage = 2
train.activation_checkpoint.enabled = False
train.num_accumulation_steps = 8
