"""GLUE dataset (reference projects/text_classification/dataset/glue_dataset.py): features are built once and cached
next to the data under a file lock."""
import logging
import os
import time
from typing import Optional, Union

import torch
from torch.utils.data import Dataset

from libai_b200.data.structures import DistTensorData, Instance
from libai_b200.utils.file_io import file_lock

from .utils import EncodePattern, Split, convert_examples_to_features
from .utils_glue import glue_output_modes, glue_processors

logger = logging.getLogger(__name__)


class _TaskDataset(Dataset):
    processors, output_modes = {}, {}

    def __init__(self, task_name, data_dir, tokenizer, max_seq_length: int = 128, mode: Union[str, Split] = Split.train,
                 pattern: Union[str, EncodePattern] = EncodePattern.bert_pattern, cache_dir: Optional[str] = None,
                 overwrite_cache: bool = True):
        self.processor = self.processors[task_name]()
        self.output_mode = self.output_modes[task_name]
        mode = Split[mode] if isinstance(mode, str) else mode
        pattern = EncodePattern[pattern] if isinstance(pattern, str) else pattern
        cached = os.path.join(cache_dir or data_dir,
                              f"cached_{mode.value}_{tokenizer.__class__.__name__}_{max_seq_length}_{task_name}")
        self.label_list = self.processor.get_labels()
        with file_lock(cached):
            if os.path.exists(cached) and not overwrite_cache:
                start = time.time()
                self.features = torch.load(cached, weights_only=False)
                logger.info(f"Loading features from cached file {cached} [took {time.time() - start:.3f} s]")
            else:
                reader = {Split.train: self.processor.get_train_examples, Split.dev: self.processor.get_dev_examples,
                          Split.test: self.processor.get_test_examples}[mode]
                self.features = convert_examples_to_features(reader(data_dir), tokenizer, max_seq_length,
                                                             self.label_list, self.output_mode, pattern)
                torch.save(self.features, cached)

    def __len__(self):
        return len(self.features)

    def __getitem__(self, i):
        f = self.features[i]
        dtype = torch.float32 if self.output_mode == "regression" else torch.long
        fields = dict(input_ids=DistTensorData(torch.tensor(f.input_ids, dtype=torch.long)),
                      attention_mask=DistTensorData(torch.tensor(f.attention_mask, dtype=torch.long)),
                      token_type_ids=DistTensorData(torch.tensor(f.token_type_ids, dtype=torch.long)))
        if f.labels is not None:
            fields["labels"] = DistTensorData(torch.tensor(f.labels, dtype=dtype), placement_idx=-1)
        return Instance(**fields)

    def get_labels(self):
        return self.label_list


class GlueDataset(_TaskDataset):
    processors, output_modes = glue_processors, glue_output_modes
