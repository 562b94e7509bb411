"""Evaluation loop and evaluator base classes.

Spec: reference libai/evaluation/evaluator.py — ``DatasetEvaluator`` (:37-80),
``DatasetEvaluators`` (:83-116), ``inference_on_dataset`` (:119-275: eval mode + no-grad, last
batch padded so every DP rank keeps the same shapes, inputs/outputs gathered to rank 0,
``evaluator.process`` on rank 0 only, data/compute/eval timing), ``inference_context`` (:279-295).
"""
from __future__ import annotations

import datetime
import logging
import time
from collections import OrderedDict, abc
from contextlib import ExitStack, contextmanager
from typing import Callable, List, Union

import torch

from libai_b200.utils import distributed as dutil
from libai_b200.utils.logger import log_every_n_seconds

from .utils import pad_batch


class DatasetEvaluator:
    """``reset()`` → many ``process(inputs, outputs)`` → ``evaluate()`` returning ``{metric: value}``."""

    def reset(self):
        pass

    def process(self, inputs, outputs):
        pass

    def evaluate(self):
        pass


class DatasetEvaluators(DatasetEvaluator):
    """Dispatch to several evaluators and merge their (disjoint) results."""

    def __init__(self, evaluators):
        super().__init__()
        self._evaluators = evaluators

    def reset(self):
        for e in self._evaluators:
            e.reset()

    def process(self, inputs, outputs):
        for e in self._evaluators:
            e.process(inputs, outputs)

    def evaluate(self):
        results = OrderedDict()
        for e in self._evaluators:
            r = e.evaluate()
            if dutil.is_main_process() and r is not None:
                for k, v in r.items():
                    assert k not in results, "Different evaluators produce results with the same key {}".format(k)
                    results[k] = v
        return results


def _gather_rows(t: torch.Tensor, valid: int) -> torch.Tensor:
    """Concatenate the first ``valid`` rows of every DP rank's tensor (0-dim tensors: DP mean)."""
    topo = dutil.get_dist_util()
    if t.dim() == 0:
        return dutil.dp_mean_to_rank0(t).cpu()
    if topo.dp_group is None:
        return t[:valid].detach().cpu()
    counts = [None] * topo.data_parallel_size
    torch.distributed.all_gather_object(counts, int(valid), group=topo.dp_group)
    full = dutil.tensor_to_rank0(t, device="cpu")
    per = t.shape[0]
    return torch.cat([full[i * per : i * per + c] for i, c in enumerate(counts)], dim=0)


def inference_on_dataset(
    model,
    data_loader,
    batch_size,
    eval_iter,
    get_batch: Callable,
    input_placement_device: str,
    evaluator: Union[DatasetEvaluator, List[DatasetEvaluator], None],
):
    """Run ``model`` over ``data_loader`` (at most ``eval_iter`` batches) and feed ``evaluator``.
    ``batch_size`` is the *global* test batch (micro × DP)."""
    logger = logging.getLogger(__name__)
    topo = dutil.get_dist_util()
    n_dev = dutil.get_world_size()
    total_samples = len(data_loader.dataset)
    if evaluator is None:
        evaluator = DatasetEvaluators([])
    if isinstance(evaluator, abc.MutableSequence):
        evaluator = DatasetEvaluators(evaluator)
    evaluator.reset()

    dps = topo.data_parallel_size
    micro = max(1, batch_size // dps)
    remain = total_samples % dps
    # ranks >= remain received one duplicated sample (index 0) in their last batch
    lack_here = 1 if (remain > 0 and topo.dp_rank >= remain) else 0
    real_iters = min(eval_iter, len(data_loader))
    n_eval = min(real_iters * batch_size, total_samples)
    logger.info(f"with eval_iter {eval_iter}, reset total samples {total_samples} to {n_eval}")
    logger.info(f"Start inference on {n_eval} samples")
    warmup = min(5, len(data_loader) - 1)
    t_start = time.perf_counter()
    t_data = t_compute = t_eval = 0.0
    consumed = 0

    with ExitStack() as stack:
        if isinstance(model, torch.nn.Module):
            stack.enter_context(inference_context(model))
        stack.enter_context(torch.no_grad())
        mark = time.perf_counter()
        for idx, inputs in enumerate(data_loader):
            if idx >= real_iters:
                break
            t_data += time.perf_counter() - mark
            if idx == warmup:
                t_start = time.perf_counter()
                t_data = t_compute = t_eval = 0.0
            c0 = time.perf_counter()
            data = get_batch(inputs, input_placement_device)
            last = idx == len(data_loader) - 1
            padded, valid = pad_batch(data, micro, lack_here if last else 0, last)
            outputs = model(**padded)
            g_in = {k: _gather_rows(v, valid) for k, v in data.items()}
            g_out = {}
            for k, v in outputs.items():
                g_out[k] = _gather_rows(v, valid) if v.dim() > 1 else _gather_rows(v if v.dim() == 0 else v, valid)
            if topo.device_type == "cuda":
                torch.cuda.synchronize()
            t_compute += time.perf_counter() - c0
            e0 = time.perf_counter()
            if dutil.is_main_process():
                evaluator.process(g_in, g_out)
            dutil.synchronize()
            t_eval += time.perf_counter() - e0
            consumed += next(iter(g_in.values())).shape[0] if g_in else valid
            done = idx + 1 - warmup * int(idx >= warmup)
            per_iter = (time.perf_counter() - t_start) / done
            if idx >= warmup * 2 or t_compute / done > 5:
                eta = datetime.timedelta(seconds=int(per_iter * (n_eval // batch_size - idx - 1)))
                log_every_n_seconds(
                    logging.INFO,
                    f"Inference done {consumed}/{n_eval}. Dataloading: {t_data / done:.4f} s/iter. "
                    f"Inference: {t_compute / done:.4f} s/iter. Eval: {t_eval / done:.4f} s/iter. "
                    f"Total: {per_iter:.4f} s/iter. ETA={eta}",
                    n=5,
                )
            mark = time.perf_counter()

    total = time.perf_counter() - t_start
    iters_timed = max(1, real_iters - warmup)
    logger.info("Total valid samples: {}".format(consumed))
    logger.info(
        "Total inference time: {} ({:.6f} s / iter per device, on {} devices)".format(
            str(datetime.timedelta(seconds=total)), total / iters_timed, n_dev
        )
    )
    logger.info(
        "Total inference pure compute time: {} ({:.6f} s / iter per device, on {} devices)".format(
            str(datetime.timedelta(seconds=int(t_compute))), t_compute / iters_timed, n_dev
        )
    )
    results = evaluator.evaluate()
    return {} if results is None else results


@contextmanager
def inference_context(model):
    """Temporarily switch ``model`` to eval mode."""
    was_training = model.training
    model.eval()
    try:
        yield
    finally:
        model.train(was_training)
