"""Evaluators: top-k accuracy, perplexity, regression correlations, BLEU; evaluator composition."""
import math

import pytest
import torch

from libai_b200.config import DictConfig
from libai_b200.evaluation import BLEUEvaluator, ClsEvaluator, PPLEvaluator, RegEvaluator
from libai_b200.evaluation.bleu_evaluator import corpus_bleu
from libai_b200.evaluation.evaluator import DatasetEvaluators
from libai_b200.utils import distributed as dist


@pytest.fixture(autouse=True)
def _topo():
    dist.setup_dist_util(DictConfig(dict(data_parallel_size=1, tensor_parallel_size=1, pipeline_parallel_size=1,
                                         device_type="cpu")))


def test_cls_evaluator_topk():
    ev = ClsEvaluator(topk=(1, 2))
    logits = torch.tensor([[0.1, 0.7, 0.2], [0.6, 0.3, 0.1], [0.2, 0.3, 0.5], [0.4, 0.35, 0.25]])
    ev.process({"labels": torch.tensor([1, 1, 2, 2])}, {"prediction_scores": logits})
    res = ev.evaluate()
    assert res["Acc@1"] == pytest.approx(50.0) and res["Acc@2"] == pytest.approx(75.0)
    ev.reset()
    assert ev.evaluate()["Acc@1"] == 0


def test_ppl_evaluator():
    ev = PPLEvaluator()
    ev.process({}, {"lm_loss": torch.tensor(1.0)})
    ev.process({}, {"lm_loss": torch.tensor(3.0)})
    assert ev.evaluate()["lm_loss_PPL"] == pytest.approx((math.e + math.e ** 3) / 2)
    ev.reset()
    ev.process({}, {"lm_loss": torch.tensor(1e6)})            # clamped: no overflow
    assert math.isfinite(ev.evaluate()["lm_loss_PPL"])


def test_reg_evaluator_correlations():
    ev = RegEvaluator()
    scores = torch.eye(5)[[0, 1, 2, 3, 4, 4]]                 # argmax = 0,1,2,3,4,4
    ev.process({"labels": torch.tensor([0, 1, 2, 3, 4, 3])}, {"prediction_scores": scores})
    res = ev.evaluate()
    assert 0.9 < res["pearson"] <= 1.0 and 0.9 < res["spearman"] <= 1.0
    assert res["corr"] == pytest.approx((res["pearson"] + res["spearman"]) / 2)


def test_bleu():
    ref = "the cat sat on the mat".split()
    assert corpus_bleu([[ref]], [ref]) == pytest.approx(1.0)
    assert corpus_bleu([[ref]], ["a dog".split()]) == 0.0
    partial = corpus_bleu([[ref]], ["the cat sat on a mat".split()])
    assert 0.3 < partial < 0.9
    short = corpus_bleu([[ref]], ["the cat sat on".split()])    # brevity penalty
    assert short < corpus_bleu([[ref]], ["the cat sat on the".split()])
    ev = BLEUEvaluator()
    ev.process({"reference": ref}, {"candidate": ref})
    assert ev.evaluate()["bleu_score"] == pytest.approx(1.0)


def test_evaluators_merge_and_reject_duplicate_keys():
    both = DatasetEvaluators([RegEvaluator(), ClsEvaluator(topk=(1,))])
    both.reset()
    scores = torch.eye(3)[[0, 1, 2, 2]]
    both.process({"labels": torch.tensor([0, 1, 2, 1])}, {"prediction_scores": scores})
    res = both.evaluate()
    assert res["Acc@1"] == pytest.approx(75.0) and "pearson" in res
    dup = DatasetEvaluators([ClsEvaluator(topk=(1,)), ClsEvaluator(topk=(1,))])
    dup.process({"labels": torch.tensor([0])}, {"prediction_scores": torch.tensor([[1.0, 0.0]])})
    with pytest.raises(AssertionError):
        dup.evaluate()
