"""Logits processors / warpers for generation.

Spec: reference libai/inference/generator/generation_logits_processor.py:25-385 — ``LogitsProcessorList`` and the 14
processors/warpers (normalisation, inf/nan removal, forced BOS/EOS, repetition penalty, Hamming diversity,
(encoder) no-repeat n-gram, min length, prefix constraint, exponential length decay, temperature, top-k, top-p,
typical).  All operate on ``scores [batch·beams, vocab]`` (fp32) given the running ``input_ids``.
"""
from __future__ import annotations

import inspect
import math
from typing import Callable, Iterable, List, Tuple

import torch


class LogitsProcessorList(list):
    def __call__(self, input_ids, scores, **kwargs):
        for processor in self:
            params = inspect.signature(processor.__call__).parameters
            if len(params) > 2:
                missing = [k for k in list(params)[2:] if k not in kwargs]
                if missing:
                    raise ValueError(f"{processor.__class__} needs {missing} to be passed to the processor list")
                scores = processor(input_ids, scores, **kwargs)
            else:
                scores = processor(input_ids, scores)
        return scores


class NormalizationLogitsProcessor:
    def __call__(self, input_ids, scores):
        return torch.log_softmax(scores, dim=-1)


class InfNanRemoveLogitsProcessor:
    def __call__(self, input_ids, scores):
        scores = torch.where(torch.isnan(scores), torch.zeros_like(scores), scores)
        return torch.where(scores == float("inf"), torch.full_like(scores, torch.finfo(scores.dtype).max), scores)


class ForcedEOSTokenLogitsProcessor:
    def __init__(self, max_length: int, eos_token_id: int):
        self.max_length, self.eos_token_id = max_length, eos_token_id

    def __call__(self, input_ids, scores):
        if input_ids.shape[-1] == self.max_length - 1:
            forced = torch.full_like(scores, -float("inf"))
            forced[:, self.eos_token_id] = 0
            return forced
        return scores


class ForcedBOSTokenLogitsProcessor:
    def __init__(self, bos_token_id: int):
        self.bos_token_id = bos_token_id

    def __call__(self, input_ids, scores):
        if input_ids.shape[-1] == 1:
            forced = torch.full_like(scores, -float("inf"))
            forced[:, self.bos_token_id] = 0
            return forced
        return scores


class RepetitionPenaltyLogitsProcessor:
    def __init__(self, penalty: float):
        if not isinstance(penalty, float) or not penalty > 0:
            raise ValueError(f"`penalty` has to be a strictly positive float, but is {penalty}")
        self.penalty = penalty

    def __call__(self, input_ids, scores):
        seen = torch.gather(scores, 1, input_ids)
        seen = torch.where(seen < 0, seen * self.penalty, seen / self.penalty)
        return scores.scatter(1, input_ids, seen)


class HammingDiversityLogitsProcessor:
    """Group beam search: penalise tokens already picked by earlier groups at this step."""

    def __init__(self, diversity_penalty: float, num_beams: int, num_beam_groups: int):
        if not isinstance(diversity_penalty, float) or not diversity_penalty > 0.0:
            raise ValueError("`diversity_penalty` should be a float strictly larger than 0.")
        if not isinstance(num_beams, int) or num_beams < 2:
            raise ValueError("`num_beams` should be an integer strictly larger than 1.")
        if not isinstance(num_beam_groups, int) or num_beam_groups < 2:
            raise ValueError("`num_beam_groups` should be an integer strictly larger than 1.")
        if num_beam_groups > num_beams:
            raise ValueError("`beam_groups` has to be smaller or equal to `num_beams`.")
        self._penalty, self._num_beams = diversity_penalty, num_beams
        self._sub_beams = num_beams // num_beam_groups

    def __call__(self, input_ids, scores, current_tokens, beam_group_idx):
        batch_size = current_tokens.shape[0] // self._num_beams
        start = beam_group_idx * self._sub_beams
        end = min(start + self._sub_beams, self._num_beams)
        size = end - start
        vocab = scores.shape[-1]
        if start == 0:
            return scores
        for b in range(batch_size):
            prev = current_tokens[b * self._num_beams : b * self._num_beams + start]
            freq = torch.bincount(prev, minlength=vocab).to(scores.dtype)
            scores[b * size : (b + 1) * size] -= self._penalty * freq
        return scores


def _banned_ngram_tokens(ngram_size: int, prev_ids: List[List[int]], cur_len: int) -> List[List[int]]:
    if cur_len + 1 < ngram_size:
        return [[] for _ in prev_ids]
    banned = []
    for seq in prev_ids:
        table = {}
        for i in range(len(seq) - ngram_size + 1):
            table.setdefault(tuple(seq[i : i + ngram_size - 1]), []).append(seq[i + ngram_size - 1])
        prefix = tuple(seq[cur_len + 1 - ngram_size : cur_len]) if ngram_size > 1 else ()
        banned.append(table.get(prefix, []))
    return banned


class NoRepeatNGramLogitsProcessor:
    def __init__(self, ngram_size: int):
        if not isinstance(ngram_size, int) or ngram_size <= 0:
            raise ValueError(f"`ngram_size` has to be a strictly positive integer, but is {ngram_size}")
        self.ngram_size = ngram_size

    def __call__(self, input_ids, scores):
        cur_len = input_ids.shape[-1]
        for i, banned in enumerate(_banned_ngram_tokens(self.ngram_size, input_ids.tolist(), cur_len)):
            if banned:
                scores[i, banned] = -float("inf")
        return scores


class EncoderNoRepeatNGramLogitsProcessor:
    """Forbid n-grams of the *encoder* input from appearing in the output."""

    def __init__(self, encoder_ngram_size: int, encoder_input_ids):
        if not isinstance(encoder_ngram_size, int) or encoder_ngram_size <= 0:
            raise ValueError("`encoder_ngram_size` has to be a strictly positive integer")
        self.ngram_size = encoder_ngram_size
        if encoder_input_ids.dim() == 1:
            encoder_input_ids = encoder_input_ids.unsqueeze(0)
        self.batch_size = encoder_input_ids.shape[0]
        self.tables = []
        for seq in encoder_input_ids.tolist():
            table = {}
            for i in range(len(seq) - encoder_ngram_size + 1):
                table.setdefault(tuple(seq[i : i + encoder_ngram_size - 1]), []).append(seq[i + encoder_ngram_size - 1])
            self.tables.append(table)

    def __call__(self, input_ids, scores):
        num_hypos = scores.shape[0]
        num_beams = num_hypos // self.batch_size
        cur_len = input_ids.shape[-1]
        for i, seq in enumerate(input_ids.tolist()):
            if cur_len + 1 < self.ngram_size:
                continue
            prefix = tuple(seq[cur_len + 1 - self.ngram_size : cur_len]) if self.ngram_size > 1 else ()
            banned = self.tables[i // num_beams].get(prefix, [])
            if banned:
                scores[i, banned] = -float("inf")
        return scores


class MinLengthLogitsProcessor:
    def __init__(self, min_length: int, eos_token_id: int):
        if not isinstance(min_length, int) or min_length < 0:
            raise ValueError(f"`min_length` has to be a positive integer, but is {min_length}")
        if not isinstance(eos_token_id, int) or eos_token_id < 0:
            raise ValueError(f"`eos_token_id` has to be a positive integer, but is {eos_token_id}")
        self.min_length, self.eos_token_id = min_length, eos_token_id

    def __call__(self, input_ids, scores):
        if input_ids.shape[-1] < self.min_length:
            scores[:, self.eos_token_id] = -float("inf")
        return scores


class PrefixConstrainedLogitsProcessor:
    def __init__(self, prefix_allowed_tokens_fn: Callable[[int, torch.Tensor], List[int]], num_beams: int):
        self._fn, self._num_beams = prefix_allowed_tokens_fn, num_beams

    def __call__(self, input_ids, scores):
        mask = torch.full_like(scores, -math.inf)
        for batch_id, beams in enumerate(input_ids.view(-1, self._num_beams, input_ids.shape[-1])):
            for beam_id, sent in enumerate(beams):
                mask[batch_id * self._num_beams + beam_id, self._fn(batch_id, sent)] = 0
        return scores + mask


class ExponentialDecayLengthPenalty:
    """After ``start_index`` generated tokens, boost EOS by ``factor ** (len - start)``."""

    def __init__(self, exponential_decay_length_penalty: Tuple, eos_token_id: int, input_ids_seq_length: int):
        self.regulation_start = exponential_decay_length_penalty[0] + input_ids_seq_length
        self.regulation_factor = exponential_decay_length_penalty[1]
        self.eos_token_id = eos_token_id

    def __call__(self, input_ids, scores):
        cur_len = input_ids.shape[-1]
        if cur_len > self.regulation_start:
            scores[:, self.eos_token_id] = scores[:, self.eos_token_id] * pow(
                self.regulation_factor, cur_len - self.regulation_start
            )
        return scores


class TemperatureLogitsWarper:
    def __init__(self, temperature: float):
        if not isinstance(temperature, float) or not temperature > 0:
            raise ValueError(f"`temperature` has to be a strictly positive float, but is {temperature}")
        self.temperature = temperature

    def __call__(self, input_ids, scores):
        return scores / self.temperature


class TopPLogitsWarper:
    def __init__(self, top_p: float, filter_value: float = -float("inf"), min_tokens_to_keep: int = 1):
        top_p = float(top_p)
        if top_p < 0 or top_p > 1.0:
            raise ValueError(f"`top_p` has to be a float > 0 and < 1, but is {top_p}")
        self.top_p, self.filter_value, self.min_tokens_to_keep = top_p, filter_value, min_tokens_to_keep

    def __call__(self, input_ids, scores):
        sorted_logits, sorted_idx = torch.sort(scores, descending=True)
        cum = sorted_logits.softmax(dim=-1).cumsum(dim=-1)
        remove = cum > self.top_p
        remove[..., 1:] = remove[..., :-1].clone()  # keep the first token that crosses the threshold
        remove[..., 0] = False
        if self.min_tokens_to_keep > 1:
            remove[..., : self.min_tokens_to_keep] = False
        remove = remove.scatter(1, sorted_idx, remove)
        return scores.masked_fill(remove, self.filter_value)


class TopKLogitsWarper:
    def __init__(self, top_k: int, filter_value: float = -float("inf"), min_tokens_to_keep: int = 1):
        if not isinstance(top_k, int) or top_k <= 0:
            raise ValueError(f"`top_k` has to be a strictly positive integer, but is {top_k}")
        self.top_k, self.filter_value, self.min_tokens_to_keep = top_k, filter_value, min_tokens_to_keep

    def __call__(self, input_ids, scores):
        k = min(max(self.top_k, self.min_tokens_to_keep), scores.shape[-1])
        kth = torch.topk(scores, k)[0][..., -1, None]
        return scores.masked_fill(scores < kth, self.filter_value)


class TypicalLogitsWarper:
    """Locally typical sampling: keep the tokens whose surprise is closest to the entropy, up to mass ``mass``."""

    def __init__(self, mass: float = 0.9, filter_value: float = -float("inf"), min_tokens_to_keep: int = 1):
        mass = float(mass)
        if not (0 < mass < 1):
            raise ValueError(f"`typical_p` has to be a float > 0 and < 1, but is {mass}")
        self.mass, self.filter_value, self.min_tokens_to_keep = mass, filter_value, min_tokens_to_keep

    def __call__(self, input_ids, scores):
        normalized = torch.log_softmax(scores, dim=-1)
        p = normalized.exp()
        ent = -torch.nan_to_num(normalized * p, nan=0.0).sum(-1, keepdim=True)
        shifted = (-normalized - ent).abs()
        sorted_scores, sorted_idx = torch.sort(shifted, descending=False)
        sorted_logits = scores.gather(-1, sorted_idx)
        cum = sorted_logits.softmax(dim=-1).cumsum(dim=-1)
        last = (cum < self.mass).sum(dim=1).clamp(max=scores.shape[-1] - 1)
        remove = sorted_scores > sorted_scores.gather(1, last.view(-1, 1))
        if self.min_tokens_to_keep > 1:
            remove[..., : self.min_tokens_to_keep] = False
        remove = remove.scatter(1, sorted_idx, remove)
        return scores.masked_fill(remove, self.filter_value)
