"""Metric storage and writers.

Spec: reference libai/utils/events.py — ``EventStorage`` (:265-450), ``JSONWriter`` (:69-135),
``TensorboardXWriter`` (:138-175), ``CommonMetricPrinter`` (:178-262, log-line format incl.
``total_throughput: X samples/s``).  Additions: the printer also reports tokens/s and the
device-timed step when the trainer records ``device_time``; TensorBoard goes through
``torch.utils.tensorboard`` (tensorboardX is not available).
"""
import datetime
import json
import logging
import os
import time
from collections import defaultdict
from contextlib import contextmanager

from .history_buffer import HistoryBuffer

__all__ = [
    "get_event_storage",
    "JSONWriter",
    "TensorboardXWriter",
    "CommonMetricPrinter",
    "EventStorage",
    "EventWriter",
]

_STORAGE_STACK = []


def get_event_storage():
    """The ``EventStorage`` of the innermost active ``with EventStorage(...)`` block."""
    assert len(_STORAGE_STACK), "get_event_storage() has to be called inside a 'with EventStorage(...)' context!"
    return _STORAGE_STACK[-1]


class EventWriter:
    def write(self):
        raise NotImplementedError

    def close(self):
        pass


class JSONWriter(EventWriter):
    """Appends one JSON object per write with the latest (smoothed) scalars → ``metrics.json``."""

    def __init__(self, json_file, window_size=20):
        os.makedirs(os.path.dirname(json_file) or ".", exist_ok=True)
        self._fh = open(json_file, "a")
        self._window = window_size
        self._last_write = -1

    def write(self):
        storage = get_event_storage()
        per_iter = defaultdict(dict)
        for k, (v, it) in storage.latest_with_smoothing_hint(self._window).items():
            if it <= self._last_write:
                continue
            per_iter[it][k] = v
        if per_iter:
            self._last_write = max(per_iter)
        for it in sorted(per_iter):
            row = per_iter[it]
            row["iteration"] = it
            self._fh.write(json.dumps(row, sort_keys=True) + "\n")
        self._fh.flush()
        try:
            os.fsync(self._fh.fileno())
        except (AttributeError, OSError):
            pass

    def close(self):
        self._fh.close()


class TensorboardXWriter(EventWriter):
    """Scalars / images / histograms to TensorBoard event files."""

    def __init__(self, log_dir: str, window_size: int = 20, **kwargs):
        self._window = window_size
        from torch.utils.tensorboard import SummaryWriter

        self._writer = SummaryWriter(log_dir=log_dir, **kwargs)
        self._last_write = -1

    def write(self):
        storage = get_event_storage()
        newest = self._last_write
        for k, (v, it) in storage.latest_with_smoothing_hint(self._window).items():
            if it > self._last_write:
                self._writer.add_scalar(k, v, it)
                newest = max(newest, it)
        self._last_write = newest
        if storage._vis_data:
            for name, img, step in storage._vis_data:
                self._writer.add_image(name, img, step)
            storage.clear_images()
        if storage._histograms:
            for params in storage._histograms:
                self._writer.add_histogram_raw(**params)
            storage.clear_histograms()

    def close(self):
        if hasattr(self, "_writer"):
            self._writer.close()


class CommonMetricPrinter(EventWriter):
    """Prints eta / iteration / consumed_samples / losses / time / throughput / lr."""

    def __init__(self, batch_size, max_iter, tokens_per_sample=None):
        self.logger = logging.getLogger(__name__)
        self._batch_size = batch_size
        self._max_iter = max_iter
        self._tokens_per_sample = tokens_per_sample
        self._last_write = None

    def write(self):
        storage = get_event_storage()
        it = storage.iter
        if it == self._max_iter:
            return  # progress printer only; nothing to say after the final iteration
        hist = storage.histories()
        data_time = hist["data_time"].avg(20) if "data_time" in hist else None
        eta = None
        iter_time = None
        if "time" in hist:
            iter_time = hist["time"].global_avg()
            eta_s = hist["time"].median(1000) * (self._max_iter - it - 1)
            storage.put_scalar("eta_seconds", eta_s, smoothing_hint=False)
            eta = str(datetime.timedelta(seconds=int(eta_s)))
        else:
            if self._last_write is not None:
                per_it = (time.perf_counter() - self._last_write[1]) / max(1, it - self._last_write[0])
                eta = str(datetime.timedelta(seconds=int(per_it * (self._max_iter - it - 1))))
            self._last_write = (it, time.perf_counter())
        lr = "{:.2e}".format(hist["lr"].latest()) if "lr" in hist else "N/A"
        max_mem = None
        try:
            import torch

            if torch.cuda.is_available():
                max_mem = torch.cuda.max_memory_allocated() / 1024.0 / 1024.0
        except Exception:
            pass
        parts = []
        if eta:
            parts.append(f"eta: {eta}")
        parts.append(f"iteration: {it}/{self._max_iter}")
        parts.append(f"consumed_samples: {storage.samples}")
        losses = "  ".join(f"{k}: {v.median(200):.4g}" for k, v in hist.items() if "loss" in k)
        if losses:
            parts.append(losses)
        if iter_time is not None:
            parts.append(f"time: {iter_time:.4f} s/iter")
        if data_time is not None:
            parts.append(f"data_time: {data_time:.4f} s/iter")
        if iter_time is not None and iter_time > 0:
            parts.append(f"total_throughput: {self._batch_size / iter_time:.2f} samples/s")
            if self._tokens_per_sample:
                parts.append(f"tokens/s: {self._batch_size * self._tokens_per_sample / iter_time:.0f}")
        if "device_time" in hist:
            parts.append(f"device_time: {hist['device_time'].median(20) * 1e3:.2f} ms")
        parts.append(f"lr: {lr}")
        if max_mem is not None:
            parts.append(f"max_mem: {max_mem:.0f}M")
        self.logger.info(" " + "  ".join(parts))


class EventStorage:
    """Scalars (with smoothing hints), images and histograms keyed by iteration."""

    def __init__(self, start_iter=0):
        self._history = defaultdict(HistoryBuffer)
        self._smoothing_hints = {}
        self._latest_scalars = {}
        self._iter = start_iter
        self._batch_size = 0
        self._samples = 0
        self._current_prefix = ""
        self._vis_data = []
        self._histograms = []

    def put_image(self, img_name, img_tensor):
        self._vis_data.append((img_name, img_tensor, self._iter))

    def put_scalar(self, name, value, smoothing_hint=True):
        name = self._current_prefix + name
        value = float(value)
        self._history[name].update(value, self._iter)
        self._latest_scalars[name] = (value, self._iter)
        prev = self._smoothing_hints.get(name)
        if prev is not None:
            assert prev == smoothing_hint, f"Scalar {name} was put with a different smoothing_hint!"
        else:
            self._smoothing_hints[name] = smoothing_hint

    def put_scalars(self, *, smoothing_hint=True, **kwargs):
        for k, v in kwargs.items():
            self.put_scalar(k, v, smoothing_hint=smoothing_hint)

    def put_histogram(self, hist_name, hist_tensor, bins=1000):
        import torch

        t = hist_tensor.detach().float().cpu()
        lo, hi = t.min().item(), t.max().item()
        counts = torch.histc(t, bins=bins)
        edges = torch.linspace(lo, hi, steps=bins + 1, dtype=torch.float32)
        self._histograms.append(
            dict(
                tag=hist_name,
                min=lo,
                max=hi,
                num=len(t),
                sum=float(t.sum()),
                sum_squares=float(torch.sum(t ** 2)),
                bucket_limits=edges[1:].tolist(),
                bucket_counts=counts.tolist(),
                global_step=self._iter,
            )
        )

    def history(self, name):
        ret = self._history.get(name)
        if ret is None:
            raise KeyError(f"No history metric available for {name}!")
        return ret

    def histories(self):
        return self._history

    def latest(self):
        return self._latest_scalars

    def latest_with_smoothing_hint(self, window_size=20):
        out = {}
        for k, (v, it) in self._latest_scalars.items():
            out[k] = (self._history[k].median(window_size) if self._smoothing_hints[k] else v, it)
        return out

    def smoothing_hints(self):
        return self._smoothing_hints

    def step(self):
        self._iter += 1

    @property
    def iter(self):
        return self._iter

    @iter.setter
    def iter(self, val):
        self._iter = int(val)

    @property
    def samples(self):
        return self._samples

    @samples.setter
    def samples(self, val):
        self._samples = int(val)

    @property
    def iteration(self):
        return self._iter

    def __enter__(self):
        _STORAGE_STACK.append(self)
        return self

    def __exit__(self, exc_type, exc_val, exc_tb):
        assert _STORAGE_STACK[-1] is self
        _STORAGE_STACK.pop()

    @contextmanager
    def name_scope(self, name):
        old = self._current_prefix
        self._current_prefix = name.rstrip("/") + "/"
        try:
            yield
        finally:
            self._current_prefix = old

    def clear_images(self):
        self._vis_data = []

    def clear_histograms(self):
        self._histograms = []
