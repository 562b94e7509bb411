"""Prompt dataset config (reference projects/MagicPrompt/configs/gpt2_dataset.py)."""
from libai_b200.config import LazyCall, OmegaConf
from libai_b200.data.build import build_nlp_test_loader, build_nlp_train_loader
from libai_b200.tokenizer import GPT2Tokenizer
from projects.MagicPrompt.datasets.datasets import PromptDataset

tokenization = OmegaConf.create()
tokenization.tokenizer = LazyCall(GPT2Tokenizer)(vocab_file="/data/home/magicprompt/vocab.json",
                                                  merges_file="/data/home/magicprompt/merges.txt")
tokenization.append_eod = False
tokenization.make_vocab_size_divisible_by = 128

dataloader = OmegaConf.create()
dataloader.train = LazyCall(build_nlp_train_loader)(
    dataset=[LazyCall(PromptDataset)(path="/data/home/magicprompt/train/en_train_mini.txt", tokenizer=tokenization.tokenizer,
                                     max_seq_length=128)],
    num_workers=4,
)
dataloader.test = [
    LazyCall(build_nlp_test_loader)(
        dataset=LazyCall(PromptDataset)(path="/data/home/magicprompt/test/en_test_mini.txt",
                                        tokenizer=tokenization.tokenizer, max_seq_length=128),
        test_batch_size=4,
    )
]
