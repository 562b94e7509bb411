"""Synthetic datasets of a named shape (benchmarks / smoke tests; there is no network for real
corpora).  Samples are deterministic functions of the index so every TP rank of a DP replica –
and the first and last pipeline stage – see identical data without communication."""
import torch
from torch.utils.data import Dataset

from libai_b200.data.structures import DistTensorData, Instance


def _gen(idx: int, seed: int) -> torch.Generator:
    g = torch.Generator()
    g.manual_seed(seed * 1000003 + idx)
    return g


class SyntheticGPTDataset(Dataset):
    """``input_ids`` / ``labels`` (next-token shifted) of length ``seq_length``."""

    def __init__(self, vocab_size=50257, seq_length=1024, num_samples=1 << 20, seed=1234):
        self.vocab_size, self.seq_length, self.num_samples, self.seed = vocab_size, seq_length, num_samples, seed

    def __len__(self):
        return self.num_samples

    def __getitem__(self, idx):
        toks = torch.randint(0, self.vocab_size, (self.seq_length + 1,), generator=_gen(idx, self.seed), dtype=torch.long)
        return Instance(
            input_ids=DistTensorData(toks[:-1].clone()),
            labels=DistTensorData(toks[1:].clone(), placement_idx=-1),
        )


class SyntheticBertDataset(Dataset):
    """BERT pre-training fields: ids, padding mask, token types, NSP label, MLM labels + loss mask."""

    def __init__(self, vocab_size=30522, seq_length=512, num_samples=1 << 20, mask_prob=0.15, seed=1234):
        self.vocab_size, self.seq_length, self.num_samples = vocab_size, seq_length, num_samples
        self.mask_prob, self.seed = mask_prob, seed

    def __len__(self):
        return self.num_samples

    def __getitem__(self, idx):
        g = _gen(idx, self.seed)
        s = self.seq_length
        ids = torch.randint(5, self.vocab_size, (s,), generator=g, dtype=torch.long)
        loss_mask = (torch.rand(s, generator=g) < self.mask_prob).long()
        labels = torch.where(loss_mask.bool(), ids, torch.full_like(ids, -1))
        half = s // 2
        return Instance(
            input_ids=DistTensorData(ids),
            attention_mask=DistTensorData(torch.ones(s, dtype=torch.long)),
            tokentype_ids=DistTensorData(torch.cat([torch.zeros(half, dtype=torch.long), torch.ones(s - half, dtype=torch.long)])),
            ns_labels=DistTensorData(torch.randint(0, 2, (1,), generator=g, dtype=torch.long).squeeze(0), placement_idx=-1),
            lm_labels=DistTensorData(labels, placement_idx=-1),
            loss_mask=DistTensorData(loss_mask, placement_idx=-1),
        )


class SyntheticImageDataset(Dataset):
    """``images`` ``[3, H, W]`` float + integer ``labels`` (ImageNet-shaped by default)."""

    def __init__(self, num_classes=1000, img_size=224, num_samples=1 << 20, seed=1234):
        self.num_classes, self.img_size, self.num_samples, self.seed = num_classes, img_size, num_samples, seed

    def __len__(self):
        return self.num_samples

    def __getitem__(self, idx):
        g = _gen(idx, self.seed)
        img = torch.randn(3, self.img_size, self.img_size, generator=g)
        label = torch.randint(0, self.num_classes, (1,), generator=g, dtype=torch.long).squeeze(0)
        return Instance(images=DistTensorData(img), labels=DistTensorData(label, placement_idx=-1))
