"""Synthetic GPT token stream (benchmarks / smoke tests, no corpus or tokenizer files needed)."""
from libai_b200.config import LazyCall, OmegaConf
from libai_b200.data import build_nlp_test_loader, build_nlp_train_loader
from libai_b200.data.datasets import SyntheticGPTDataset

dataloader = OmegaConf.create()
dataloader.train = LazyCall(build_nlp_train_loader)(
    dataset=[LazyCall(SyntheticGPTDataset)(vocab_size=50304, seq_length=1024, num_samples=1 << 20, seed=1234)],
    num_workers=2,
)
dataloader.test = [
    LazyCall(build_nlp_test_loader)(
        dataset=LazyCall(SyntheticGPTDataset)(vocab_size=50304, seq_length=1024, num_samples=64, seed=4321),
        test_batch_size=4,
        num_workers=0,
    )
]
