"""CLUE task processors (reference projects/text_classification/dataset/utils_clue.py): json-lines releases."""
import os

from .utils import DataProcessor, InputExample


class _Json(DataProcessor):
    a_key, b_key, label_key, labels = "sentence", None, "label", ["0", "1"]

    def _examples(self, rows, set_type):
        return [InputExample(f"{set_type}-{i}", r[self.a_key], r.get(self.b_key) if self.b_key else None,
                             str(r[self.label_key]) if set_type != "test" and self.label_key in r else None)
                for i, r in enumerate(rows)]

    def get_train_examples(self, data_dir):
        return self._examples(self._read_json(os.path.join(data_dir, "train.json")), "train")

    def get_dev_examples(self, data_dir):
        return self._examples(self._read_json(os.path.join(data_dir, "dev.json")), "dev")

    def get_test_examples(self, data_dir):
        return self._examples(self._read_json(os.path.join(data_dir, "test.json")), "test")

    def get_labels(self):
        return self.labels


class AfqmcProcessor(_Json):
    a_key, b_key = "sentence1", "sentence2"


class TnewsProcessor(_Json):
    labels = [str(100 + i) for i in range(17) if i not in (5, 11)]


class IflytekProcessor(_Json):
    labels = [str(i) for i in range(119)]


class OcnliProcessor(_Json):
    a_key, b_key, labels = "sentence1", "sentence2", ["contradiction", "entailment", "neutral"]

    def _examples(self, rows, set_type):
        return super()._examples([r for r in rows if r.get("label", "x") != "-"], set_type)


class CmnliProcessor(OcnliProcessor):
    pass


class CslProcessor(_Json):
    def _examples(self, rows, set_type):
        return [InputExample(f"{set_type}-{i}", " ".join(r["keyword"]), r["abst"],
                             str(r["label"]) if set_type != "test" else None) for i, r in enumerate(rows)]


class WscProcessor(_Json):
    labels = ["true", "false"]

    def _examples(self, rows, set_type):
        out = []
        for i, r in enumerate(rows):
            text = list(r["text"])
            t = r["target"]
            qi, pi = t["span1_index"], t["span2_index"]
            q, p = t["span1_text"], t["span2_text"]
            if pi > qi:
                text.insert(qi, "_"); text.insert(qi + len(q) + 1, "_")
                text.insert(pi + 2, "["); text.insert(pi + len(p) + 3, "]")
            else:
                text.insert(pi, "["); text.insert(pi + len(p) + 1, "]")
                text.insert(qi + 2, "_"); text.insert(qi + len(q) + 3, "_")
            out.append(InputExample(f"{set_type}-{i}", "".join(text), None, str(r["label"]) if set_type != "test" else None))
        return out


class CopaProcessor(_Json):
    """COPA (CLUE version): every record yields TWO sentence-pair examples, one per choice — for an "effect" question
    (premise, choice), for a "cause" question (choice, premise).  Labels follow the reference verbatim (reference
    utils_clue.py:397-441: both examples of a record carry ``1 if label == 0 else 0``)."""

    def _examples(self, rows, set_type):
        out = []
        for i, r in enumerate(rows):
            label = None if set_type == "test" else str(1 if r["label"] == 0 else 0)
            for j, choice in enumerate((r["choice0"], r["choice1"])):
                if r["question"] == "effect":
                    a, b = r["premise"], choice
                elif r["question"] == "cause":
                    a, b = choice, r["premise"]
                else:
                    raise ValueError(f"unknown question type {r['question']!r}")
                out.append(InputExample(f"{set_type}-{2 * i + j}", a, b, label))
        return out


clue_processors = {"copa": CopaProcessor, "afqmc": AfqmcProcessor, "tnews": TnewsProcessor, "iflytek": IflytekProcessor, "cmnli": CmnliProcessor,
                   "ocnli": OcnliProcessor, "csl": CslProcessor, "wsc": WscProcessor}
clue_output_modes = {k: "classification" for k in clue_processors}
clue_tasks_num_labels = {k: len(v().get_labels()) for k, v in clue_processors.items()}
