"""SimCSE corpora (reference projects/SimCSE/dataset/dataset.py): SNLI / STS / LCQMC / wiki loaders, sentence-pair
(unsup: the sentence twice; sup: triples) tokenisation with fixed-length padding."""
import json
import random

import torch
from torch.utils.data import Dataset

from libai_b200.data.structures import DistTensorData, Instance


def load_data(name, path):
    def rows(fn):
        with open(path, "r", encoding="utf-8") as f:
            return [fn(ln) for ln in f if ln.strip()]

    if name == "snli-unsup":
        return rows(lambda ln: json.loads(ln).get("origin", json.loads(ln).get("sentence1")))
    if name == "snli-sup":
        return rows(lambda ln: tuple(json.loads(ln)[k] for k in ("origin", "entailment", "contradiction")))
    if name == "lqcmc":
        return rows(lambda ln: ln.rstrip("\n").split("\t")[0])
    if name == "cnsd_sts":
        return rows(lambda ln: (ln.split("||")[1], ln.split("||")[2], ln.split("||")[3].strip()))
    if name == "wiki":
        return rows(lambda ln: ln.strip())
    if name == "eng_sts":
        return rows(lambda ln: (ln.split("\t")[5], ln.split("\t")[6].strip(), ln.split("\t")[4]))
    if name == "sts_to_train":
        return rows(lambda ln: ln.split("||")[1]) + rows(lambda ln: ln.split("||")[2])
    raise ValueError(name)


def padding_for_ids(ids, pad_id=0, max_len=64):
    ids = ids[:max_len]
    mask = [1] * len(ids) + [0] * (max_len - len(ids))
    return ids + [pad_id] * (max_len - len(ids)), mask


class _Base(Dataset):
    def __init__(self, tokenizer, max_len):
        self.tokenizer, self.max_len = tokenizer, max_len
        self.cls, self.sep, self.pad = tokenizer.cls_token_id, tokenizer.sep_token_id, tokenizer.pad_token_id

    def encode(self, text):
        ids = self.tokenizer.convert_tokens_to_ids(self.tokenizer.tokenize(text))[: self.max_len - 2]
        return padding_for_ids([self.cls] + ids + [self.sep], self.pad, self.max_len)

    def pack(self, texts, label=None):
        enc = [self.encode(t) for t in texts]
        fields = dict(input_ids=DistTensorData(torch.tensor([e[0] for e in enc], dtype=torch.long)),
                      attention_mask=DistTensorData(torch.tensor([e[1] for e in enc], dtype=torch.long)))
        if label is not None:
            fields["labels"] = DistTensorData(torch.tensor(int(float(label)), dtype=torch.long), placement_idx=-1)
        return Instance(**fields)

    def __len__(self):
        return len(self.data)


class TrainDataset_unsup(_Base):
    def __init__(self, name, path, tokenizer, max_len, path2=None):
        super().__init__(tokenizer, max_len)
        self.data = load_data(name, path) + (load_data("sts_to_train", path2) if path2 else [])
        random.shuffle(self.data)

    def __getitem__(self, index):
        return self.pack([self.data[index], self.data[index]])


class TestDataset_unsup(_Base):
    def __init__(self, name, path, tokenizer, max_len=64):
        super().__init__(tokenizer, max_len)
        self.data = load_data(name, path)

    def __getitem__(self, index):
        a, b, label = self.data[index]
        return self.pack([a, b], label)


class TrainDataset_sup(_Base):
    def __init__(self, name, path, tokenizer, max_len=64):
        super().__init__(tokenizer, max_len)
        self.data = load_data(name, path)

    def __getitem__(self, index):
        return self.pack(list(self.data[index]))


TestDataset_sup = TestDataset_unsup
