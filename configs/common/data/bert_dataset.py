"""BertDataset over a Megatron-format indexed corpus (``<prefix>.bin/.idx``); one corpus is split into
train/valid/test by ``splits`` (keys follow reference configs/common/data/bert_dataset.py)."""
from libai_b200.config import LazyCall, OmegaConf
from libai_b200.data import build_nlp_test_loader, build_nlp_train_val_test_loader
from libai_b200.data.data_utils import get_indexed_dataset
from libai_b200.data.datasets import BertDataset
from libai_b200.tokenizer import BertTokenizer

DATA_PREFIX = "/workspace/data/libai_dataset/loss_compara_content_sentence"

tokenization = OmegaConf.create()
tokenization.tokenizer = LazyCall(BertTokenizer)(
    vocab_file="bert-base-chinese-vocab.txt",
    do_lower_case=True,
    do_chinese_wwm=True,
)
tokenization.append_eod = False
tokenization.make_vocab_size_divisible_by = 128


def _corpus(**extra):
    return LazyCall(BertDataset)(
        name="bert",
        data_prefix=DATA_PREFIX,
        indexed_dataset=LazyCall(get_indexed_dataset)(data_prefix=DATA_PREFIX, data_impl="mmap", skip_warmup=False),
        max_seq_length=512,
        mask_lm_prob=0.15,
        short_seq_prob=0.1,
        binary_head=True,
        seed=1234,
        masking_style='bert-cn-wwm',
        **extra,
    )


dataloader = OmegaConf.create()
dataloader.train = LazyCall(build_nlp_train_val_test_loader)(
    dataset=[_corpus()],
    train_val_test_num_samples=None,  # filled in by the trainer from the iteration counts
    splits=[[949.0, 50.0, 1.0]],
    weights=[1.0],
    num_workers=4,
)
dataloader.test = [
    LazyCall(build_nlp_test_loader)(dataset=_corpus(max_num_samples=10), test_batch_size=4),
]
