"""Perplexity = exp(loss) averaged over evaluation batches (spec: reference
libai/evaluation/ppl_evaluator.py:25-60; the model is called *with labels* and returns losses)."""
import copy
import math
from collections import OrderedDict

from libai_b200.utils import distributed as dutil

from .evaluator import DatasetEvaluator


class PPLEvaluator(DatasetEvaluator):
    def __init__(self):
        self._predictions = []

    def reset(self):
        self._predictions = []

    def process(self, inputs, outputs):
        for k, v in outputs.items():
            self._predictions.append({f"{k}_PPL": math.exp(min(20, float(v)))})

    def evaluate(self):
        if not dutil.is_main_process():
            return {}
        sums, n = OrderedDict(), max(len(self._predictions), 1)
        for pred in self._predictions:
            for k, v in pred.items():
                sums[k] = sums.get(k, 0.0) + v
        self._results = OrderedDict((k, v / n) for k, v in sums.items())
        return copy.deepcopy(self._results)
