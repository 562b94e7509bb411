"""QQP fine-tuning (reference projects/QQP/configs/config_qqp.py)."""
from libai_b200.config import LazyCall, OmegaConf, get_config
from libai_b200.data.build import build_nlp_test_loader, build_nlp_train_loader
from libai_b200.evaluation import ClsEvaluator
from projects.QQP.dataset.qqp_dataset import QQPDataset
from projects.QQP.modeling.model import Classification
from projects.QQP.tokenizer.tokenizer import _BertCNWWMTokenizer

tokenization = get_config("common/data/bert_dataset.py").tokenization
optim = get_config("common/optim.py").optim
model_cfg = get_config("common/models/bert.py").cfg
graph = get_config("common/models/graph.py").graph
train = get_config("common/train.py").train

tokenization.tokenizer = LazyCall(_BertCNWWMTokenizer)(vocab_file="projects/QQP/QQP_DATA/bert-base-chinese-vocab.txt", lower_case=True)
tokenization.append_eod = False
tokenization.make_vocab_size_divisible_by = 128

dataloader = OmegaConf.create()
dataloader.train = LazyCall(build_nlp_train_loader)(
    dataset=[LazyCall(QQPDataset)(dataset_name="QQP_TRAIN", data_paths=["projects/QQP/QQP_DATA/train.tsv"],
                                  tokenizer=tokenization.tokenizer, max_seq_length=512)],
    num_workers=4,
)
dataloader.test = [
    LazyCall(build_nlp_test_loader)(
        dataset=LazyCall(QQPDataset)(dataset_name="QQP_TEST", data_paths=["projects/QQP/QQP_DATA/dev.tsv"],
                                     tokenizer=tokenization.tokenizer, max_seq_length=512),
        num_workers=4,
    ),
]

model_cfg.update(dict(vocab_size=21248, hidden_size=1024, hidden_layers=24, num_attention_heads=16, intermediate_size=4096,
                      num_classes=2, pretrain_megatron_weight=None))
model = LazyCall(Classification)(cfg=model_cfg)

optim.lr = 1e-5
train.update(
    dict(
        activation_checkpoint=dict(enabled=True), output_dir="output/finetune_qqp/", train_micro_batch_size=16,
        test_micro_batch_size=4, train_epoch=1, train_iter=0,
        evaluation=dict(enabled=True, eval_period=500, evaluator=LazyCall(ClsEvaluator)(topk=(1,))),
        log_period=50, dist=dict(data_parallel_size=1, tensor_parallel_size=1, pipeline_parallel_size=1),
    )
)
