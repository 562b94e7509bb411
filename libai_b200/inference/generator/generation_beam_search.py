"""Beam search bookkeeping.

Spec: reference libai/inference/generator/generation_beam_search.py — ``BeamScorer`` (:28-38), ``BeamHypotheses``
(:41-87; n-best heap with length-penalised score ``sum_logprobs / len**length_penalty``), ``BeamSearchScorer``
(:90-336; ``process`` picks the next ``num_beams`` continuations out of ``2·num_beams`` candidates and retires
finished ones, ``finalize`` pads/returns the best ``num_beam_hyps_to_keep`` per batch item).
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import List, Optional, Tuple

import torch


class BeamScorer(ABC):
    @abstractmethod
    def process(self, input_ids, next_scores, next_tokens, next_indices, **kwargs):
        raise NotImplementedError("This is an abstract method.")

    @abstractmethod
    def finalize(self, input_ids, next_scores, next_tokens, next_indices, max_length, **kwargs):
        raise NotImplementedError("This is an abstract method.")


class BeamHypotheses:
    def __init__(self, num_beams: int, length_penalty: float, early_stopping: bool):
        self.length_penalty, self.early_stopping, self.num_beams = length_penalty, early_stopping, num_beams
        self.beams: List[Tuple[float, torch.Tensor, Optional[torch.Tensor]]] = []
        self.worst_score = 1e9

    def __len__(self):
        return len(self.beams)

    def add(self, hyp: torch.Tensor, sum_logprobs: float, beam_indices=None):
        score = sum_logprobs / (hyp.shape[-1] ** self.length_penalty)
        if len(self) < self.num_beams or score > self.worst_score:
            self.beams.append((score, hyp, beam_indices))
            if len(self) > self.num_beams:
                ranked = sorted((s, i) for i, (s, _, _) in enumerate(self.beams))
                del self.beams[ranked[0][1]]
                self.worst_score = ranked[1][0]
            else:
                self.worst_score = min(score, self.worst_score)

    def is_done(self, best_sum_logprobs: float, cur_len: int) -> bool:
        """No open beam can still beat the worst kept hypothesis."""
        if len(self) < self.num_beams:
            return False
        if self.early_stopping:
            return True
        return self.worst_score >= best_sum_logprobs / cur_len ** self.length_penalty


class BeamSearchScorer(BeamScorer):
    def __init__(self, batch_size: int, num_beams: int, length_penalty: Optional[float] = 1.0,
                 do_early_stopping: Optional[bool] = False, num_beam_hyps_to_keep: Optional[int] = 1,
                 num_beam_groups: Optional[int] = 1, **kwargs):
        self.num_beams, self.length_penalty, self.do_early_stopping = num_beams, length_penalty, do_early_stopping
        self.num_beam_hyps_to_keep, self.num_beam_groups = num_beam_hyps_to_keep, num_beam_groups
        self.group_size = num_beams // num_beam_groups
        self._beam_hyps = [BeamHypotheses(num_beams, length_penalty, do_early_stopping) for _ in range(batch_size)]
        self._done = torch.zeros(batch_size, dtype=torch.bool)
        if not isinstance(num_beams, int) or num_beams <= 1:
            raise ValueError(f"`num_beams` has to be an integer strictly greater than 1, but is {num_beams}.")
        if not isinstance(num_beam_groups, int) or num_beam_groups > num_beams or num_beams % num_beam_groups != 0:
            raise ValueError("`num_beam_groups` has to be an integer smaller or equal than `num_beams` and divide it.")

    @property
    def is_done(self) -> bool:
        return bool(self._done.all())

    def process(self, input_ids, next_scores, next_tokens, next_indices, pad_token_id=None, eos_token_id=None,
                beam_indices=None):
        cur_len = input_ids.shape[-1]
        batch_size = len(self._beam_hyps)
        if batch_size != input_ids.shape[0] // self.group_size:
            raise ValueError("A group beam size does not match the number of input rows")
        device = input_ids.device
        out_scores = torch.zeros((batch_size, self.group_size), dtype=next_scores.dtype, device=device)
        out_tokens = torch.zeros((batch_size, self.group_size), dtype=next_tokens.dtype, device=device)
        out_indices = torch.zeros((batch_size, self.group_size), dtype=next_indices.dtype, device=device)
        nt, ns, ni = next_tokens.tolist(), next_scores.tolist(), next_indices.tolist()
        for b, hyps in enumerate(self._beam_hyps):
            if self._done[b]:
                if pad_token_id is None:
                    raise ValueError("A finished batch item needs `pad_token_id` to pad its beams")
                out_tokens[b, :] = pad_token_id
                continue
            slot = 0
            for rank, (tok, score, idx) in enumerate(zip(nt[b], ns[b], ni[b])):
                row = b * self.group_size + idx
                if eos_token_id is not None and tok == eos_token_id:
                    if rank >= self.group_size:  # an EOS outside the top `num_beams` never becomes a hypothesis
                        continue
                    hyps.add(input_ids[row].clone(), score,
                             beam_indices=None if beam_indices is None else beam_indices[row] + (row,))
                else:
                    out_scores[b, slot], out_tokens[b, slot], out_indices[b, slot] = score, tok, row
                    slot += 1
                if slot == self.group_size:
                    break
            if slot < self.group_size:
                raise ValueError(f"At most {self.group_size} tokens can be EOS among the 2·beams candidates")
            self._done[b] = self._done[b] or hyps.is_done(max(ns[b]), cur_len)
        return {"next_beam_scores": out_scores.view(-1), "next_beam_tokens": out_tokens.view(-1),
                "next_beam_indices": out_indices.view(-1)}

    def finalize(self, input_ids, final_beam_scores, final_beam_tokens=None, final_beam_indices=None, max_length=None,
                 pad_token_id=None, eos_token_id=None, beam_indices=None):
        batch_size = len(self._beam_hyps)
        for b, hyps in enumerate(self._beam_hyps):
            if self._done[b]:
                continue
            for beam in range(self.num_beams):  # open beams become hypotheses
                row = b * self.num_beams + beam
                hyps.add(input_ids[row], final_beam_scores[row].item(),
                         beam_indices=None if beam_indices is None else beam_indices[row])
        keep = self.num_beam_hyps_to_keep
        best, best_scores, lengths = [], torch.zeros(batch_size * keep, dtype=torch.float32), []
        for b, hyps in enumerate(self._beam_hyps):
            ranked = sorted(hyps.beams, key=lambda x: x[0])
            for j in range(keep):
                score, hyp, _ = ranked.pop()
                best.append(hyp)
                lengths.append(len(hyp))
                best_scores[b * keep + j] = score
        sent_max_len = min(max(lengths) + 1, max_length) if max_length is not None else max(lengths) + 1
        decoded = input_ids.new_zeros((batch_size * keep, sent_max_len))
        if min(lengths) != max(lengths) or sent_max_len > max(lengths):
            assert pad_token_id is not None, "`pad_token_id` has to be defined"
            decoded.fill_(pad_token_id)
        for i, hyp in enumerate(best):
            decoded[i, : lengths[i]] = hyp
            if lengths[i] < sent_max_len and eos_token_id is not None:
                decoded[i, lengths[i]] = eos_token_id
        return {"sequences": decoded, "sequence_scores": best_scores}
