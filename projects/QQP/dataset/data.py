"""GLUE dataset base class of the QQP example (reference projects/QQP/dataset/data.py)."""
import logging
from abc import ABC, abstractmethod

import torch
from torch.utils.data import Dataset

from libai_b200.data.structures import DistTensorData, Instance

from .data_utils import build_sample, build_tokens_types_paddings_from_text

logger = logging.getLogger(__name__)


class GLUEAbstractDataset(ABC, Dataset):
    def __init__(self, task_name, dataset_name, datapaths, tokenizer, max_seq_length):
        self.task_name, self.dataset_name = task_name, dataset_name
        self.tokenizer, self.max_seq_length = tokenizer, max_seq_length
        self.samples = []
        for path in ([datapaths] if isinstance(datapaths, str) else datapaths):
            self.samples.extend(self.process_samples_from_single_path(path))
        logger.info(f"  >> total number of samples: {len(self.samples)}")

    def __len__(self):
        return len(self.samples)

    def __getitem__(self, idx):
        raw = self.samples[idx]
        ids, types, paddings = build_tokens_types_paddings_from_text(raw["text_a"], raw["text_b"], self.tokenizer, self.max_seq_length)
        s = build_sample(ids, types, paddings, raw["label"], raw["uid"])
        return Instance(
            model_input=DistTensorData(torch.from_numpy(s["text"])),
            attention_mask=DistTensorData(torch.from_numpy(s["padding_mask"])),
            tokentype_ids=DistTensorData(torch.from_numpy(s["types"])),
            labels=DistTensorData(torch.tensor(s["label"], dtype=torch.long), placement_idx=-1),
        )

    @abstractmethod
    def process_samples_from_single_path(self, datapath):
        """list of ``{"text_a", "text_b", "label", "uid"}``"""
