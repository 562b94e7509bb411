"""Spearman correlation between predicted cosine similarity and gold STS scores (reference projects/SimCSE/evaluator.py)."""
from collections import OrderedDict

from libai_b200.evaluation.evaluator import DatasetEvaluator
from libai_b200.utils import distributed as dist


def spearman_target(cos_sim, labels):
    from scipy.stats import spearmanr

    return spearmanr(labels, cos_sim).correlation


class SimcseEvaluator(DatasetEvaluator):
    def __init__(self):
        self._predictions = []

    def reset(self):
        self._predictions = []

    def process(self, inputs, outputs):
        self._predictions.append({"sim": outputs["sim"].float().cpu(), "labels": outputs["labels"].float().cpu()})

    def evaluate(self):
        if not dist.is_main_process():
            return {}
        sim = [x for p in self._predictions for x in p["sim"].reshape(-1).tolist()]
        labels = [x for p in self._predictions for x in p["labels"].reshape(-1).tolist()]
        return OrderedDict(Spearman=spearman_target(sim, labels))
