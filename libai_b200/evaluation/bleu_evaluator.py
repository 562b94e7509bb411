"""Corpus BLEU (spec: reference libai/evaluation/bleu_evaluator.py:26-64, which calls
``nltk.translate.bleu_score.corpus_bleu``).  nltk is not available, so corpus BLEU-4 with uniform
weights and brevity penalty is implemented here (identical definition, no smoothing)."""
import copy
import math
from collections import Counter, OrderedDict

from libai_b200.utils import distributed as dutil

from .evaluator import DatasetEvaluator


def _ngrams(tokens, n):
    return Counter(tuple(tokens[i : i + n]) for i in range(len(tokens) - n + 1))


def corpus_bleu(list_of_references, hypotheses, max_n=4):
    """``list_of_references[i]`` is a list of reference token lists for ``hypotheses[i]``."""
    num = [0] * max_n
    den = [0] * max_n
    hyp_len = ref_len = 0
    for refs, hyp in zip(list_of_references, hypotheses):
        if refs and not isinstance(refs[0], (list, tuple)):
            refs = [refs]
        hyp_len += len(hyp)
        ref_len += min((abs(len(r) - len(hyp)), len(r)) for r in refs)[1]
        for n in range(1, max_n + 1):
            h = _ngrams(hyp, n)
            max_ref = Counter()
            for r in refs:
                for g, c in _ngrams(r, n).items():
                    max_ref[g] = max(max_ref[g], c)
            num[n - 1] += sum(min(c, max_ref[g]) for g, c in h.items())
            den[n - 1] += max(sum(h.values()), 0)
    if hyp_len == 0 or num[0] == 0:
        return 0.0
    log_p = 0.0
    for n_, d_ in zip(num, den):
        if n_ == 0 or d_ == 0:
            return 0.0
        log_p += math.log(n_ / d_) / max_n
    bp = 1.0 if hyp_len > ref_len else math.exp(1 - ref_len / hyp_len)
    return bp * math.exp(log_p)


class BLEUEvaluator(DatasetEvaluator):
    def __init__(self):
        super().__init__()
        self._predictions = []

    def reset(self):
        self._predictions = []

    def process(self, inputs, outputs):
        self._predictions.append({"candidate": outputs["candidate"], "reference": inputs["reference"]})

    def evaluate(self):
        if not dutil.is_main_process():
            return {}
        cands = [p["candidate"] for p in self._predictions]
        refs = [p["reference"] for p in self._predictions]
        self._results = OrderedDict(bleu_score=corpus_bleu(refs, cands))
        return copy.deepcopy(self._results)
