"""Pearson / Spearman correlation for regression-style tasks (spec: reference
libai/evaluation/reg_evaluator.py:30-68)."""
import copy
from collections import OrderedDict

import numpy as np

from libai_b200.utils import distributed as dutil

from .evaluator import DatasetEvaluator


class RegEvaluator(DatasetEvaluator):
    def __init__(self):
        self._predictions = []

    def reset(self):
        self._predictions = []

    def process(self, inputs, outputs):
        scores, labels = outputs["prediction_scores"], inputs["labels"]
        preds = scores.float().cpu().topk(1)[1].squeeze(1).numpy()
        self._predictions.append({"preds": preds, "labels": labels.cpu().numpy()})

    def evaluate(self):
        if not dutil.is_main_process():
            return {}
        from scipy.stats import pearsonr, spearmanr

        preds = np.concatenate([p["preds"] for p in self._predictions]) if self._predictions else np.array([])
        labels = np.concatenate([p["labels"] for p in self._predictions]) if self._predictions else np.array([])
        pearson = pearsonr(preds, labels)[0]
        spearman = spearmanr(preds, labels)[0]
        self._results = OrderedDict(pearson=pearson, spearman=spearman, corr=(pearson + spearman) / 2)
        return copy.deepcopy(self._results)
