"""Top-k classification accuracy in percent (spec: reference libai/evaluation/cls_evaluator.py:24-84)."""
import copy
from collections import OrderedDict

from libai_b200.utils import distributed as dutil

from .evaluator import DatasetEvaluator


def accuracy(output, target, topk=(1,)):
    """Percentage of rows whose label is among the ``k`` highest scores, for each ``k``."""
    maxk = min(max(topk), output.size(1))
    n = target.size(0)
    ranked = output.topk(maxk, dim=1, largest=True, sorted=True).indices
    hits = ranked.eq(target.reshape(-1, 1))
    return [hits[:, : min(k, maxk)].any(dim=1).float().sum().item() * 100.0 / n for k in topk]


class ClsEvaluator(DatasetEvaluator):
    def __init__(self, topk=(1, 5)):
        self.topk = topk
        self._predictions = []

    def reset(self):
        self._predictions = []

    def process(self, inputs, outputs):
        logits, labels = outputs["prediction_scores"], inputs["labels"]
        accs = accuracy(logits.float(), labels, topk=self.topk)
        n = labels.size(0)
        self._predictions.append({"num_correct_topk": [a * n / 100 for a in accs], "num_samples": n})

    def evaluate(self):
        if not dutil.is_main_process():
            return {}
        correct = OrderedDict(("Acc@" + str(k), 0) for k in self.topk)
        total = 0
        for pred in self._predictions:
            for k, c in zip(self.topk, pred["num_correct_topk"]):
                correct["Acc@" + str(k)] += int(round(c))
            total += int(pred["num_samples"])
        self._results = OrderedDict((k, v / max(total, 1) * 100) for k, v in correct.items())
        return copy.deepcopy(self._results)
