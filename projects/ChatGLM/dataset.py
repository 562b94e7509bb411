"""SFT dataset (reference projects/ChatGLM/dataset.py): ``prompt`` / ``response`` jsonl → ``[gMASK] sop prompt
response eos`` with the prompt part of the labels masked (−100), padded to ``max_source_len + max_target_len``."""
import json

import torch
from torch.utils.data import Dataset

from libai_b200.data.structures import DistTensorData, Instance

IGNORE_INDEX = -100


class ChatGLMTrainDataset(Dataset):
    def __init__(self, path, tokenizer, max_source_len=128, max_target_len=128, max_length=None):
        self.tokenizer = tokenizer
        self.max_len = max_length or (max_source_len + max_target_len)
        self.max_source_len, self.max_target_len = max_source_len, max_target_len
        with open(path, "r", encoding="utf-8") as f:
            text = f.read().strip()
        self.data = json.loads(text) if text.startswith("[") else [json.loads(ln) for ln in text.splitlines() if ln.strip()]

    def __len__(self):
        return len(self.data)

    def __getitem__(self, index):
        item = self.data[index]
        prompt = item.get("prompt", item.get("instruction", "") + item.get("input", ""))
        response = item.get("response", item.get("output", ""))
        a = self.tokenizer.get_prefix_tokens() + self.tokenizer.tokenizer.encode(prompt)[: self.max_source_len - 2]
        b = self.tokenizer.tokenizer.encode(response)[: self.max_target_len - 1] + [self.tokenizer.eos_token_id]
        ids = (a + b)[: self.max_len]
        labels = ([IGNORE_INDEX] * len(a) + b)[: self.max_len]
        pad = self.max_len - len(ids)
        ids = ids + [self.tokenizer.pad_token_id] * pad
        labels = labels + [IGNORE_INDEX] * pad
        return Instance(input_ids=DistTensorData(torch.tensor(ids, dtype=torch.long)),
                        labels=DistTensorData(torch.tensor(labels, dtype=torch.long), placement_idx=-1))
