"""GLUE task processors (reference projects/text_classification/dataset/utils_glue.py): file names, columns and label
sets of the official tsv releases."""
from .utils import DataProcessor, InputExample


class _Tsv(DataProcessor):
    train_file, dev_file, test_file = "train.tsv", "dev.tsv", "test.tsv"
    skip_header = True
    cols = (0, None, 1)          # (text_a, text_b, label) column indices for train/dev
    test_cols = (1, None)        # (text_a, text_b) for test
    labels = ["0", "1"]

    def _examples(self, lines, set_type):
        out = []
        for i, line in enumerate(lines[1 if self.skip_header else 0:]):
            if set_type == "test":
                a, b = self.test_cols
                out.append(InputExample(f"{set_type}-{i}", line[a], line[b] if b is not None else None, None))
            else:
                a, b, c = self.cols
                out.append(InputExample(f"{set_type}-{i}", line[a], line[b] if b is not None else None, line[c]))
        return out

    def get_train_examples(self, data_dir):
        import os

        return self._examples(self._read_tsv(os.path.join(data_dir, self.train_file)), "train")

    def get_dev_examples(self, data_dir):
        import os

        return self._examples(self._read_tsv(os.path.join(data_dir, self.dev_file)), "dev")

    def get_test_examples(self, data_dir):
        import os

        return self._examples(self._read_tsv(os.path.join(data_dir, self.test_file)), "test")

    def get_labels(self):
        return self.labels


class ColaProcessor(_Tsv):
    skip_header, cols, test_cols = False, (3, None, 1), (1, None)

    def _examples(self, lines, set_type):
        if set_type == "test":
            lines = lines[1:]
        return super()._examples(lines, set_type)


class Sst2Processor(_Tsv):
    cols, test_cols = (0, None, 1), (1, None)


class MrpcProcessor(_Tsv):
    cols, test_cols = (3, 4, 0), (3, 4)


class StsbProcessor(_Tsv):
    cols, test_cols, labels = (7, 8, 9), (7, 8), [None]


class QqpProcessor(_Tsv):
    cols, test_cols = (3, 4, 5), (1, 2)

    def _examples(self, lines, set_type):
        lines = [ln for i, ln in enumerate(lines) if i == 0 or len(ln) > (2 if set_type == "test" else 5)]
        return super()._examples(lines, set_type)


class MnliProcessor(_Tsv):
    dev_file, test_file = "dev_matched.tsv", "test_matched.tsv"
    cols, test_cols, labels = (8, 9, -1), (8, 9), ["contradiction", "entailment", "neutral"]


class MnliMismatchedProcessor(MnliProcessor):
    dev_file, test_file = "dev_mismatched.tsv", "test_mismatched.tsv"


class QnliProcessor(_Tsv):
    cols, test_cols, labels = (1, 2, -1), (1, 2), ["entailment", "not_entailment"]


class RteProcessor(QnliProcessor):
    pass


class WnliProcessor(_Tsv):
    cols, test_cols = (1, 2, -1), (1, 2)


glue_processors = {
    "cola": ColaProcessor, "mnli": MnliProcessor, "mnli-mm": MnliMismatchedProcessor, "mrpc": MrpcProcessor,
    "sst-2": Sst2Processor, "sts-b": StsbProcessor, "qqp": QqpProcessor, "qnli": QnliProcessor, "rte": RteProcessor,
    "wnli": WnliProcessor,
}
glue_output_modes = {k: ("regression" if k == "sts-b" else "classification") for k in glue_processors}
glue_tasks_num_labels = {k: (1 if k == "sts-b" else len(v().get_labels())) for k, v in glue_processors.items()}
