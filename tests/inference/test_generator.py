"""Generator: greedy / sampling / beam search, cache consistency, processors, stopping criteria."""
import math
import time

import pytest
import torch

from libai_b200.config import DictConfig
from libai_b200.inference.generator import (
    BeamSearchScorer,
    LogitsProcessorList,
    MaxLengthCriteria,
    MaxTimeCriteria,
    StoppingCriteriaList,
)
from libai_b200.inference.generator.generation_logits_processor import (
    ForcedBOSTokenLogitsProcessor,
    ForcedEOSTokenLogitsProcessor,
    MinLengthLogitsProcessor,
    NoRepeatNGramLogitsProcessor,
    RepetitionPenaltyLogitsProcessor,
    TemperatureLogitsWarper,
    TopKLogitsWarper,
    TopPLogitsWarper,
    TypicalLogitsWarper,
)
from libai_b200.models import LlamaForCausalLM


@pytest.fixture(scope="module")
def model():
    torch.manual_seed(0)
    cfg = DictConfig(dict(
        hidden_layers=2, vocab_size=100, hidden_size=64, intermediate_size=128, num_attention_heads=4,
        max_position_embeddings=64, rms_norm_eps=1e-5, initializer_range=0.02, use_scaled_init_for_output_weights=True,
        scale_mask_softmax_fusion=False, amp_enabled=False, bos_token_id=1, eos_token_id=2, pad_token_id=0, max_length=20))
    from libai_b200.utils import distributed as dutil

    dutil.reset_dist_util()
    return LlamaForCausalLM(cfg).eval()


def test_greedy_cache_equals_no_cache_and_bruteforce(model):
    ids = torch.randint(3, 100, (2, 5))
    a = model.generate(ids, max_length=12)
    b = model.generate(ids, max_length=12, use_cache=False)
    assert torch.equal(a, b)
    cur = ids.clone()
    with torch.no_grad():
        for _ in range(7):
            cur = torch.cat([cur, model(cur)["logits"][:, -1].argmax(-1, keepdim=True)], 1)
    done = (a == 2).cumsum(1) > 0
    assert torch.equal(cur[~done], a[~done])


def test_sampling_and_beam_shapes(model):
    ids = torch.randint(3, 100, (2, 5))
    torch.manual_seed(1)
    s = model.generate(ids, max_length=10, do_sample=True, top_k=5, top_p=0.9, temperature=0.8, num_return_sequences=2)
    assert s.shape == (4, 10) and torch.equal(s[:, :5], ids.repeat_interleave(2, 0))
    b = model.generate(ids, max_length=10, num_beams=3, num_return_sequences=2, no_repeat_ngram_size=2)
    assert b.shape[0] == 4 and b.shape[1] <= 10
    seqs, scores, steps = model.generate(ids, max_length=9, num_beams=2, output_scores=True)
    assert seqs.shape[0] == 2 and scores.shape == (2,) and len(steps) == 4


def test_beam_search_beats_or_equals_greedy_logprob(model):
    ids = torch.randint(3, 100, (1, 4))

    def logprob(seq):
        with torch.no_grad():
            lp = torch.log_softmax(model(seq[:, :-1])["logits"].float(), -1)
        tgt = seq[:, 1:]
        return lp.gather(-1, tgt[..., None]).squeeze(-1)[:, ids.shape[1] - 1 :].sum().item()

    g = model.generate(ids, max_length=9, eos_token_id=None, pad_token_id=0)
    b = model.generate(ids, max_length=9, num_beams=4, eos_token_id=None, pad_token_id=0, length_penalty=0.0)
    assert logprob(b[:, :9]) >= logprob(g) - 1e-4


def test_logits_processors():
    ids = torch.tensor([[1, 2, 3, 1, 2]])
    scores = torch.zeros(1, 6)
    assert NoRepeatNGramLogitsProcessor(3)(ids, scores.clone())[0, 3] == -math.inf          # "1 2 3" seen
    assert MinLengthLogitsProcessor(10, 4)(ids, scores.clone())[0, 4] == -math.inf
    assert ForcedEOSTokenLogitsProcessor(6, 5)(ids, scores.clone())[0].argmax() == 5
    assert ForcedBOSTokenLogitsProcessor(2)(ids[:, :1], scores.clone())[0].argmax() == 2
    rp = RepetitionPenaltyLogitsProcessor(2.0)(ids, torch.tensor([[1.0, 1.0, -1.0, 1.0, 1.0, 1.0]]))
    assert rp[0].tolist() == [1.0, 0.5, -2.0, 0.5, 1.0, 1.0]
    s = torch.tensor([[1.0, 2.0, 3.0, 4.0]])
    assert (TopKLogitsWarper(2)(ids, s.clone()) == -math.inf).sum() == 2
    assert torch.allclose(TemperatureLogitsWarper(2.0)(ids, s.clone()), s / 2)
    p = TopPLogitsWarper(0.6)(ids, torch.log(torch.tensor([[0.1, 0.2, 0.3, 0.4]])))
    assert (p == -math.inf).tolist() == [[True, True, False, False]]
    t = TypicalLogitsWarper(0.5)(ids, torch.log(torch.tensor([[0.1, 0.2, 0.3, 0.4]])))
    assert (t > -math.inf).sum() >= 1
    lst = LogitsProcessorList([MinLengthLogitsProcessor(10, 4), TemperatureLogitsWarper(2.0)])
    assert lst(ids, scores.clone())[0, 4] == -math.inf


def test_stopping_criteria():
    ids = torch.zeros(1, 5, dtype=torch.long)
    assert MaxLengthCriteria(5)(ids, None) and not MaxLengthCriteria(6)(ids, None)
    assert MaxTimeCriteria(0.0, initial_timestamp=time.time() - 1)(ids, None)
    lst = StoppingCriteriaList([MaxLengthCriteria(7)])
    assert lst.max_length == 7 and not lst(ids, None)


def test_beam_scorer_finalize_prefers_better_hypothesis():
    scorer = BeamSearchScorer(batch_size=1, num_beams=2, length_penalty=1.0)
    ids = torch.tensor([[5, 6], [5, 7]])
    out = scorer.process(ids, torch.tensor([[-0.1, -0.5, -2.0, -3.0]]), torch.tensor([[9, 2, 8, 7]]),
                         torch.tensor([[0, 0, 1, 1]]), pad_token_id=0, eos_token_id=2)
    assert out["next_beam_tokens"].tolist() == [9, 8] and out["next_beam_indices"].tolist() == [0, 1]
    fin = scorer.finalize(torch.tensor([[5, 6, 9], [5, 7, 8]]), out["next_beam_scores"], pad_token_id=0, eos_token_id=2,
                          max_length=5)
    assert fin["sequences"].shape[0] == 1
