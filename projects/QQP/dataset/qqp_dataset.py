"""Quora Question Pairs (reference projects/QQP/dataset/qqp_dataset.py): train/dev rows have 6 columns
(id, qid1, qid2, q1, q2, is_duplicate), test rows 3 (id, q1, q2)."""
import logging

from .data import GLUEAbstractDataset
from .data_utils import clean_text

logger = logging.getLogger(__name__)
LABELS = [0, 1]


class QQPDataset(GLUEAbstractDataset):
    def __init__(self, dataset_name, data_paths, tokenizer, max_seq_length, test_label=0):
        self.test_label = test_label
        super().__init__("QQP", dataset_name, data_paths, tokenizer, max_seq_length)

    def process_samples_from_single_path(self, filename):
        logger.info(f" > Processing {filename} ...")
        samples, is_test, first = [], False, True
        with open(filename, "r", encoding="utf-8") as f:
            for line in f:
                row = line.strip().split("\t")
                if first:
                    first = False
                    is_test = len(row) == 3
                    continue
                if is_test:
                    assert len(row) == 3, f"expected length 3: {row}"
                    uid, a, b, label = int(row[0].strip()), clean_text(row[1].strip()), clean_text(row[2].strip()), self.test_label
                elif len(row) == 6:
                    uid, a, b, label = int(row[0].strip()), clean_text(row[3].strip()), clean_text(row[4].strip()), int(row[5].strip())
                else:
                    logger.info(f"***WARNING*** index error, skipping: {row}")
                    continue
                if not a or not b:
                    continue
                assert label in LABELS and uid >= 0
                samples.append({"uid": uid, "text_a": a, "text_b": b, "label": label})
        return samples
