"""Couplet pairs (reference projects/Couplets/dataset/dataset.py): ``{train,test}/in.txt`` holds the first lines,
``out.txt`` the matching second lines (space separated characters), plus ``vocabs``."""
import os

import torch
from torch.utils.data import Dataset

from libai_b200.data.structures import DistTensorData, Instance
from projects.Couplets.dataset.mask import make_padding_mask, make_sequence_mask
from projects.Couplets.tokenizer.tokenizer import CoupletsTokenizer


class CoupletsDataset(Dataset):
    def __init__(self, path, is_train=True, maxlen=64):
        split = "train" if is_train else "test"
        with open(os.path.join(path, split, "in.txt"), encoding="utf-8") as f:
            self.src = [ln.strip() for ln in f if ln.strip()]
        with open(os.path.join(path, split, "out.txt"), encoding="utf-8") as f:
            self.tgt = [ln.strip() for ln in f if ln.strip()]
        assert len(self.src) == len(self.tgt)
        self.tokenizer = CoupletsTokenizer(os.path.join(path, "vocabs"))
        self.maxlen = maxlen

    def __len__(self):
        return len(self.src)

    def text2ids(self, text):
        t = self.tokenizer
        ids = [t.bos_id] + t.convert_tokens_to_ids(t.tokenize(text))[: self.maxlen - 2] + [t.eos_id]
        return ids + [t.pad_id] * (self.maxlen - len(ids))

    def __getitem__(self, index):
        pad = self.tokenizer.pad_id
        enc, full = self.text2ids(self.src[index]), self.text2ids(self.tgt[index])
        dec, labels = full[:-1] + [pad], full[1:] + [pad]
        return Instance(
            encoder_input_ids=DistTensorData(torch.tensor(enc, dtype=torch.long)),
            decoder_input_ids=DistTensorData(torch.tensor(dec, dtype=torch.long)),
            encoder_attn_mask=DistTensorData(torch.from_numpy(make_padding_mask(enc, enc, pad))),
            decoder_attn_mask=DistTensorData(torch.from_numpy(make_padding_mask(dec, dec, pad) * make_sequence_mask(dec))),
            encoder_decoder_attn_mask=DistTensorData(torch.from_numpy(make_padding_mask(dec, enc, pad))),
            lm_labels=DistTensorData(torch.tensor(labels, dtype=torch.long), placement_idx=-1),
        )
