"""Example / feature containers and the tsv/json readers shared by the GLUE and CLUE processors (reference
projects/text_classification/dataset/utils.py)."""
import csv
import json
from dataclasses import dataclass
from enum import Enum
from typing import List, Optional, Union


class Split(Enum):
    train = "train"
    dev = "dev"
    test = "test"


class EncodePattern(Enum):
    bert_pattern = "bert_pattern"        # [CLS] A [SEP] B [SEP]
    roberta_pattern = "roberta_pattern"  # [CLS] A [SEP] [SEP] B [SEP]


@dataclass
class InputExample:
    guid: str
    text_a: str
    text_b: Optional[str] = None
    label: Optional[str] = None


@dataclass(frozen=True)
class InputFeatures:
    input_ids: List[int]
    attention_mask: Optional[List[int]] = None
    token_type_ids: Optional[List[int]] = None
    labels: Optional[Union[int, float]] = None


class DataProcessor:
    """Per-task reader: ``get_{train,dev,test}_examples(data_dir)`` and ``get_labels()``."""

    def get_train_examples(self, data_dir):
        raise NotImplementedError

    def get_dev_examples(self, data_dir):
        raise NotImplementedError

    def get_test_examples(self, data_dir):
        raise NotImplementedError

    def get_labels(self):
        raise NotImplementedError

    @classmethod
    def _read_tsv(cls, input_file, quotechar=None):
        with open(input_file, "r", encoding="utf-8-sig") as f:
            return list(csv.reader(f, delimiter="\t", quotechar=quotechar))

    @classmethod
    def _read_json(cls, input_file):
        with open(input_file, "r", encoding="utf-8") as f:
            return [json.loads(ln) for ln in f if ln.strip()]


def _truncate_pair(a, b, max_length):
    a, b = list(a), list(b)
    while len(a) + len(b) > max_length:
        (a if len(a) > len(b) else b).pop()
    return a, b


def convert_examples_to_features(examples, tokenizer, max_length, label_list=None, output_mode="classification",
                                 pattern=EncodePattern.bert_pattern):
    label_map = {label: i for i, label in enumerate(label_list or [])}
    cls, sep, pad = tokenizer.cls_token_id, tokenizer.sep_token_id, tokenizer.pad_token_id or 0
    features = []
    for ex in examples:
        a = tokenizer.convert_tokens_to_ids(tokenizer.tokenize(ex.text_a))
        b = tokenizer.convert_tokens_to_ids(tokenizer.tokenize(ex.text_b)) if ex.text_b else []
        extra = 3 if pattern == EncodePattern.bert_pattern else 4
        if b:
            a, b = _truncate_pair(a, b, max_length - extra)
            mid = [sep] if pattern == EncodePattern.bert_pattern else [sep, sep]
            ids = [cls] + a + mid + b + [sep]
            types = [0] * (len(a) + 1 + len(mid)) + [1] * (len(b) + 1)
            if pattern == EncodePattern.roberta_pattern:
                types = [0] * len(ids)
        else:
            a = a[: max_length - 2]
            ids, types = [cls] + a + [sep], [0] * (len(a) + 2)
        mask = [1] * len(ids)
        padn = max_length - len(ids)
        ids, mask, types = ids + [pad] * padn, mask + [0] * padn, types + [0] * padn
        label = None
        if ex.label is not None:
            label = float(ex.label) if output_mode == "regression" else label_map[ex.label]
        features.append(InputFeatures(ids, mask, types, label))
    return features
