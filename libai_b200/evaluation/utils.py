"""Evaluation helpers (spec: reference libai/evaluation/utils.py:24-93)."""
import logging
from collections.abc import Mapping

import torch


def pad_batch(x_dict, batch_size, last_batch_lack=0, is_last_batch=False, device=None):
    """Zero-pad every tensor of a *local* batch to ``batch_size`` rows so all DP ranks run the same
    shapes on the last (short) batch.  Returns ``(padded_dict, valid_rows)`` where ``valid_rows``
    excludes the ``last_batch_lack`` duplicate samples the sampler appended on this rank."""
    first = next(iter(x_dict.values()))
    n = first.shape[0]
    assert n <= batch_size
    if n == batch_size and not is_last_batch:
        return x_dict, batch_size
    valid = n - last_batch_lack
    if n == batch_size:
        return x_dict, valid
    out = {}
    for k, x in x_dict.items():
        pad = torch.zeros((batch_size,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        pad[:n] = x
        out[k] = pad
    return out, valid


def print_csv_format(results):
    """Log the main metrics in a copy-paste friendly form (``task -> {metric: score}``)."""
    assert isinstance(results, Mapping) or not len(results), results
    logger = logging.getLogger(__name__)
    for task, res in results.items():
        if isinstance(res, Mapping):
            main = [(k, v) for k, v in res.items() if "-" not in k]
            logger.info("copypaste: Task: {}".format(task))
            logger.info("copypaste: " + ",".join(k for k, _ in main))
            logger.info("copypaste: " + ",".join("{0:.4f}".format(v) for _, v in main))
        else:
            logger.info(f"copypaste: {task}={res}")


def flatten_results_dict(results):
    """``{"a": {"b": {"c": v}}}`` → ``{"a/b/c": v}``."""
    flat = {}
    for k, v in results.items():
        if isinstance(v, Mapping):
            for kk, vv in flatten_results_dict(v).items():
                flat[f"{k}/{kk}"] = vv
        else:
            flat[k] = v
    return flat
