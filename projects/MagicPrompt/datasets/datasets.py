"""Prompt corpus for MagicPrompt fine-tuning (reference projects/MagicPrompt/datasets/datasets.py): one prompt per
line, tokenised with GPT-2 BPE, concatenated with ``<|endoftext|>`` and cut into fixed-length training blocks."""
import torch
from torch.utils.data import Dataset

from libai_b200.data.structures import DistTensorData, Instance


class PromptDataset(Dataset):
    def __init__(self, path, tokenizer, max_seq_length=128):
        self.tokenizer, self.max_seq_length = tokenizer, max_seq_length
        with open(path, "r", encoding="utf-8") as f:
            lines = [ln.strip() for ln in f if ln.strip()]
        eos = tokenizer.eos_token_id
        stream = []
        for ln in lines:
            stream.extend(tokenizer.encode(ln) + [eos])
        n = (len(stream) - 1) // max_seq_length
        self.blocks = torch.tensor(stream[: n * max_seq_length + 1], dtype=torch.long)
        self.n = n

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        chunk = self.blocks[i * self.max_seq_length : (i + 1) * self.max_seq_length + 1]
        return Instance(input_ids=DistTensorData(chunk[:-1].clone()),
                        labels=DistTensorData(chunk[1:].clone(), placement_idx=-1))
