"""Pausable wall-clock timer (spec: reference libai/utils/timer.py:26-86) plus a CUDA-event
device timer used for the official device-timed throughput metric."""
from time import perf_counter
from typing import Optional


class Timer:
    """Measures elapsed ``perf_counter`` time excluding paused intervals."""

    def __init__(self):
        self.reset()

    def reset(self):
        self._t0 = perf_counter()
        self._pause_t0: Optional[float] = None
        self._paused_total = 0.0
        self._n_resume = 1

    def pause(self):
        if self._pause_t0 is not None:
            raise ValueError("Trying to pause a Timer that is already paused!")
        self._pause_t0 = perf_counter()

    def is_paused(self) -> bool:
        return self._pause_t0 is not None

    def resume(self):
        if self._pause_t0 is None:
            raise ValueError("Trying to resume a Timer that is not paused!")
        self._paused_total += perf_counter() - self._pause_t0
        self._pause_t0 = None
        self._n_resume += 1

    def seconds(self) -> float:
        end = self._pause_t0 if self._pause_t0 is not None else perf_counter()
        return end - self._t0 - self._paused_total

    def avg_seconds(self) -> float:
        return self.seconds() / self._n_resume


class DeviceTimer:
    """CUDA-event timer on the current stream; ``elapsed_ms`` synchronises on the stop event."""

    def __init__(self):
        import torch

        self._torch = torch
        self._start = torch.cuda.Event(enable_timing=True)
        self._stop = torch.cuda.Event(enable_timing=True)

    def start(self):
        self._start.record()

    def stop(self):
        self._stop.record()

    def elapsed_ms(self) -> float:
        self._stop.synchronize()
        return self._start.elapsed_time(self._stop)
