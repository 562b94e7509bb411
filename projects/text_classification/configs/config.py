"""CLUE AFQMC fine-tuning of BERT-large (reference projects/text_classification/configs/config.py)."""
from libai_b200.config import LazyCall, OmegaConf, get_config
from libai_b200.data.build import build_nlp_test_loader, build_nlp_train_loader
from libai_b200.evaluation import ClsEvaluator
from libai_b200.tokenizer import BertTokenizer
from projects.text_classification.dataset import ClueDataset
from projects.text_classification.modeling.model import ModelForSequenceClassification

tokenization = get_config("common/data/bert_dataset.py").tokenization
optim = get_config("common/optim.py").optim
model_cfg = get_config("common/models/bert.py").cfg
graph = get_config("common/models/graph.py").graph
train = get_config("common/train.py").train

tokenization.tokenizer = LazyCall(BertTokenizer)(vocab_file="./bert-base-chinese-vocab.txt", do_lower_case=True,
                                                  do_chinese_wwm=False)
tokenization.append_eod = False
tokenization.make_vocab_size_divisible_by = 128

_data_dir = "./projects/text_classification/dataset/clue_data/afqmc"
dataloader = OmegaConf.create()
dataloader.train = LazyCall(build_nlp_train_loader)(
    dataset=[LazyCall(ClueDataset)(task_name="afqmc", data_dir=_data_dir, tokenizer=tokenization.tokenizer,
                                   max_seq_length=128, mode="train")],
    num_workers=4,
)
dataloader.test = [
    LazyCall(build_nlp_test_loader)(
        dataset=LazyCall(ClueDataset)(task_name="afqmc", data_dir=_data_dir, tokenizer=tokenization.tokenizer,
                                      max_seq_length=512, mode="dev"),
        num_workers=4,
    ),
]

model_cfg.update(dict(vocab_size=21248, hidden_size=1024, hidden_layers=24, num_attention_heads=16, num_classes=2,
                      pretrain_megatron_weight=None))
model = LazyCall(ModelForSequenceClassification)(cfg=model_cfg)

train.update(
    dict(
        activation_checkpoint=dict(enabled=True), output_dir="output/benchmark/", train_micro_batch_size=4,
        test_micro_batch_size=4, train_epoch=1, train_iter=0,
        evaluation=dict(enabled=True, eval_period=500, evaluator=LazyCall(ClsEvaluator)(topk=(1,))),
        log_period=50, dist=dict(data_parallel_size=1, tensor_parallel_size=1, pipeline_parallel_size=1),
    )
)
