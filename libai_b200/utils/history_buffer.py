"""Scalar time series with windowed statistics (spec: reference libai/utils/history_buffer.py:26-89)."""
from typing import List, Optional, Tuple

import numpy as np


class HistoryBuffer:
    def __init__(self, max_length: int = 1000000):
        self._max_length = max_length
        self._data: List[Tuple[float, float]] = []  # (value, iteration)
        self._count = 0
        self._global_sum = 0.0

    def update(self, value: float, iteration: Optional[float] = None) -> None:
        if iteration is None:
            iteration = self._count
        if len(self._data) == self._max_length:
            self._data.pop(0)
        self._data.append((value, iteration))
        self._count += 1
        self._global_sum += value

    def latest(self) -> float:
        return self._data[-1][0]

    def median(self, window_size: int) -> float:
        return float(np.median([v for v, _ in self._data[-window_size:]]))

    def avg(self, window_size: int) -> float:
        return float(np.mean([v for v, _ in self._data[-window_size:]]))

    def global_avg(self) -> float:
        return self._global_sum / self._count

    def values(self) -> List[Tuple[float, float]]:
        return self._data
