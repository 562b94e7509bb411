"""Stopping criteria (reference libai/inference/generator/generation_stopping_criteria.py:25-68)."""
import time
import warnings
from copy import deepcopy
from typing import Optional


class StoppingCriteriaList(list):
    def __call__(self, input_ids, scores, **kwargs) -> bool:
        return any(criteria(input_ids, scores) for criteria in self)

    @property
    def max_length(self) -> Optional[int]:
        for criteria in self:
            if isinstance(criteria, MaxLengthCriteria):
                return criteria.max_length
        return None


class MaxLengthCriteria:
    def __init__(self, max_length: int):
        self.max_length = max_length

    def __call__(self, input_ids, scores) -> bool:
        return input_ids.shape[-1] >= self.max_length


class MaxTimeCriteria:
    def __init__(self, max_time: float, initial_timestamp: Optional[float] = None):
        self.max_time = max_time
        self.initial_timestamp = time.time() if initial_timestamp is None else initial_timestamp

    def __call__(self, input_ids, scores) -> bool:
        return time.time() - self.initial_timestamp > self.max_time


def validate_stopping_criteria(stopping_criteria: StoppingCriteriaList, max_length: int) -> StoppingCriteriaList:
    stopping_max_length = stopping_criteria.max_length
    new = deepcopy(stopping_criteria)
    if stopping_max_length is not None and stopping_max_length != max_length:
        warnings.warn("You set different `max_length` for stopping criteria and `max_length` parameter", UserWarning)
    elif stopping_max_length is None:
        new.append(MaxLengthCriteria(max_length=max_length))
    return new
