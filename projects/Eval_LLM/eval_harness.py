"""Language-model evaluation harness adapter.

Spec: reference projects/Eval_LLM/eval_harness.py:23-350 — wrap a causal LM + tokenizer behind the three
lm-evaluation-harness request types (``loglikelihood``, ``loglikelihood_rolling``, ``generate_until``), batch the
token-level scoring by length, and run a task list.

Differences by design:
* the scoring core has no dependency on ``lm_eval``: the request types are implemented here, so the adapter runs
  (and is unit-tested) without the package.  When ``lm_eval`` is importable, ``as_lm_eval_model()`` wraps the same
  object in an ``lm_eval.api.model.LM`` subclass and ``run_eval`` drives the upstream evaluator;
* task files can also be local JSONL (``LocalTask``) — multiple choice (acc / acc_norm), perplexity, and
  generate-until exact match — which covers the offline case (no dataset download);
* the continuation log-probabilities are gathered from a fused ``log_softmax`` over only the continuation rows
  rather than the whole ``[B, S, V]`` tensor: at vocab 150k+ the full fp32 log-softmax would be the largest
  allocation of the run.
"""
from __future__ import annotations

import fnmatch
import json
import math
import os
from pathlib import Path
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import torch

from libai_b200.utils import distributed as dist


# --------------------------------------------------------------------------------------------- token windows
def get_rolling_token_windows(token_list: Sequence[int], prefix_token: int, max_seq_len: int, context_len: int = 1):
    """Split a long document into (context, prediction) windows so every token is predicted exactly once, each with
    at least ``context_len`` tokens of context and at most ``max_seq_len`` tokens per model call."""
    assert 1 <= context_len <= max_seq_len
    if not token_list:
        return
    pred_len = max_seq_len - context_len + 1
    predicted = 0
    first = min(max_seq_len, len(token_list))
    yield [prefix_token] + list(token_list[: first - 1]), list(token_list[:first])
    predicted += first
    while predicted < len(token_list):
        window_pred_len = min(len(token_list) - predicted, pred_len)
        window_end = predicted + window_pred_len
        yield (
            list(token_list[window_end - max_seq_len - 1: window_end - 1]),
            list(token_list[window_end - window_pred_len: window_end]),
        )
        predicted += window_pred_len


def make_disjoint_window(pair):
    """(context, prediction) with overlap → context that stops where the prediction starts."""
    a, b = pair
    return a[: len(a) - (len(b) - 1)], b


def _args(request):
    for attr in ("arguments", "args"):
        if hasattr(request, attr):
            return getattr(request, attr)
    return request


# --------------------------------------------------------------------------------------------- the adapter
class EvalHarnessBase:
    def __init__(self, model, tokenizer, model_name: str, batch_size: int, cfg=None):
        self.model = model
        self.tokenizer = tokenizer
        self.model_name = model_name
        self.batch_size_per_gpu = int(batch_size)
        self.cfg = cfg if cfg is not None else getattr(model, "cfg", None)

    # ---- properties the harness queries
    def _cfg_get(self, key, default=None):
        cfg = self.cfg
        if cfg is None:
            return default
        val = cfg.get(key, None) if hasattr(cfg, "get") else getattr(cfg, key, None)
        return default if val is None else val

    @property
    def eos_token_id(self):
        tid = getattr(self.tokenizer, "eos_token_id", None)
        return tid if tid is not None else self._cfg_get("eos_token_id")

    eot_token_id = eos_token_id

    @property
    def pad_token_id(self):
        tid = getattr(self.tokenizer, "pad_token_id", None)
        if tid is None:
            tid = self._cfg_get("pad_token_id")
        return tid if tid is not None else self.eos_token_id

    @property
    def max_length(self):
        return int(self._cfg_get("max_position_embeddings", 1024))

    @property
    def vocab_size(self):
        return int(self._cfg_get("vocab_size"))

    @property
    def max_gen_toks(self):
        return int(self._cfg_get("max_length", 256))

    @property
    def batch_size(self):
        return self.batch_size_per_gpu

    @property
    def device(self):
        try:
            return next(self.model.parameters()).device
        except StopIteration:
            return torch.device("cpu")

    # ---- tokenizer access (HF tokenizers and this repo's tokenizers both work)
    def tok_encode(self, string: str) -> List[int]:
        tok = self.tokenizer
        if hasattr(tok, "encode"):
            try:
                out = tok.encode(string, add_special_tokens=False)
            except TypeError:
                out = tok.encode(string)
        else:
            out = tok.convert_tokens_to_ids(tok.tokenize(string))
        return [int(t) for t in (out.tolist() if hasattr(out, "tolist") else out)]

    def tok_decode(self, tokens: Iterable[int]) -> str:
        return self.tokenizer.decode([int(t) for t in tokens])

    # ---- model access
    @torch.inference_mode()
    def _model_call(self, inps: torch.Tensor) -> torch.Tensor:
        out = self.model(inps.to(self.device))
        return out["logits"] if isinstance(out, dict) else out

    @torch.inference_mode()
    def _model_generate(self, context: torch.Tensor, max_length: int, eos_token_id) -> torch.Tensor:
        return self.model.generate(context.to(self.device), max_length=max_length, eos_token_id=eos_token_id,
                                   pad_token_id=self.pad_token_id, do_sample=False)

    # ---- request types
    def loglikelihood(self, requests, disable_tqdm=True) -> List[Tuple[float, bool]]:
        new_reqs = []
        for request in requests:
            context, continuation = _args(request)
            context_enc = [self.eos_token_id] if context == "" else self.tok_encode(context)
            continuation_enc = self.tok_encode(continuation)[: self.max_length]
            new_reqs.append(((context, continuation), context_enc, continuation_enc))
        return self._loglikelihood_tokens(new_reqs)

    def loglikelihood_rolling(self, requests) -> List[float]:
        out = []
        for request in requests:
            (string,) = _args(request)
            windows = [
                (None,) + make_disjoint_window(w)
                for w in get_rolling_token_windows(self.tok_encode(string), self.eos_token_id, self.max_length, 1)
            ]
            out.append(sum(x[0] for x in self._loglikelihood_tokens(windows)))
        return out

    def _loglikelihood_tokens(self, requests, disable_tqdm=True) -> List[Tuple[float, bool]]:
        """requests: (key, context_tokens, continuation_tokens).  Longest first, so the first batch sets the
        high-water mark of activation memory and an OOM shows up immediately rather than an hour in."""
        order = sorted(range(len(requests)), key=lambda i: -(len(requests[i][1]) + len(requests[i][2])))
        results: List[Optional[Tuple[float, bool]]] = [None] * len(requests)
        bs = max(1, self.batch_size_per_gpu)
        for start in range(0, len(order), bs):
            chunk = [requests[i] for i in order[start: start + bs]]
            rows, meta = [], []
            for _, context_enc, continuation_enc in chunk:
                assert len(continuation_enc) > 0, "empty continuation"
                # the model sees everything but the last continuation token, left-truncated to max_length
                inp = (list(context_enc) + list(continuation_enc))[-(self.max_length + 1):][:-1]
                rows.append(inp)
                meta.append((len(inp), len(continuation_enc)))
            width = max(len(r) for r in rows)
            batch = torch.full((len(rows), width), int(self.pad_token_id), dtype=torch.long)
            for r, row in enumerate(rows):
                batch[r, : len(row)] = torch.tensor(row, dtype=torch.long)   # right padding: causal → harmless
            logits = self._model_call(batch)
            for r, ((inplen, contlen), (_, _, continuation_enc)) in enumerate(zip(meta, chunk)):
                contlen = min(contlen, inplen)
                cont_logits = logits[r, inplen - contlen: inplen].float()
                logp = torch.log_softmax(cont_logits, dim=-1)
                target = torch.tensor(list(continuation_enc)[-contlen:], dtype=torch.long, device=logp.device)
                greedy = bool((logp.argmax(dim=-1) == target).all().item())
                score = float(logp.gather(-1, target[:, None]).sum().item())
                results[order[start + r]] = (score, greedy)
        return results

    def generate_until(self, requests, disable_tqdm=True) -> List[str]:
        out = []
        for request in requests:
            context, gen_kwargs = _args(request)
            gen_kwargs = dict(gen_kwargs or {})
            until = gen_kwargs.get("until", None)
            until = [until] if isinstance(until, str) else list(until or [])
            max_gen_toks = int(gen_kwargs.get("max_gen_toks", self.max_gen_toks))
            ctx = self.tok_encode(context)[-(self.max_length - max_gen_toks):]
            ids = torch.tensor([ctx], dtype=torch.long)
            gen = self._model_generate(ids, len(ctx) + max_gen_toks, self.eos_token_id)
            text = self.tok_decode(gen[0, len(ctx):].tolist())
            eos_text = self.tok_decode([self.eos_token_id]) if self.eos_token_id is not None else None
            for stop in until + ([eos_text] if eos_text else []):
                if stop:
                    text = text.split(stop)[0]
            out.append(text)
        return out

    # ---- lm_eval bridge
    def as_lm_eval_model(self):
        from lm_eval.api.model import LM  # noqa: deferred, optional dependency

        base = self

        class _Bridge(LM):
            def __init__(self):
                super().__init__()

            loglikelihood = staticmethod(base.loglikelihood)
            loglikelihood_rolling = staticmethod(base.loglikelihood_rolling)
            generate_until = staticmethod(base.generate_until)

        return _Bridge()

    @torch.inference_mode()
    def run_eval(self, eval_tasks: List[str], limit: Optional[int] = None, bootstrap_iters: int = 100000) -> Dict:
        local = [t for t in eval_tasks if str(t).endswith(".jsonl") or os.path.exists(str(t))]
        named = [t for t in eval_tasks if t not in local]
        results = {"results": {}}
        for path in local:
            task = LocalTask(path)
            results["results"][task.name] = task.evaluate(self, limit=limit)
        if named:
            try:
                from lm_eval import evaluator, tasks
            except ImportError as e:
                raise ImportError(
                    f"tasks {named} are lm-evaluation-harness task names but `lm_eval` is not installed; "
                    "install it or pass local *.jsonl task files"
                ) from e
            manager = tasks.TaskManager()
            names = sorted({m for pat in named for m in fnmatch.filter(manager.all_tasks, pat)})
            print(f"Found tasks: {names}")
            if dist.is_main_process():
                tasks.get_task_dict(names)       # download/cache once
            dist.synchronize()
            up = evaluator.evaluate(lm=self.as_lm_eval_model(), task_dict=tasks.get_task_dict(names), limit=limit,
                                    bootstrap_iters=bootstrap_iters)
            results["results"].update(up["results"])
        results["config"] = dict(model=self.model_name, batch_size=self.batch_size, device=str(self.device),
                                 limit=limit, bootstrap_iters=bootstrap_iters)
        return results


# --------------------------------------------------------------------------------------------- local tasks
class LocalTask:
    """A JSONL task file; each line is one document of one of three kinds:

    * ``{"query": str, "choices": [str, ...], "gold": int}`` → ``acc`` and length-normalised ``acc_norm``
    * ``{"text": str}`` → ``word_perplexity`` / ``byte_perplexity`` / ``bits_per_byte``
    * ``{"query": str, "until": [str, ...], "answer": str}`` → ``exact_match``
    """

    def __init__(self, path):
        self.path = Path(path)
        self.name = self.path.stem
        with open(self.path, "r", encoding="utf-8") as f:
            self.docs = [json.loads(line) for line in f if line.strip()]

    def evaluate(self, lm: EvalHarnessBase, limit: Optional[int] = None) -> Dict[str, float]:
        docs = self.docs[:limit] if limit else self.docs
        mc = [d for d in docs if "choices" in d]
        ppl = [d for d in docs if "text" in d]
        gen = [d for d in docs if "answer" in d and "choices" not in d]
        metrics: Dict[str, float] = {}
        if mc:
            reqs = [(d["query"], c) for d in mc for c in d["choices"]]
            scores = lm.loglikelihood(reqs)
            k, acc, acc_norm = 0, 0, 0
            for d in mc:
                ll = [scores[k + i][0] for i in range(len(d["choices"]))]
                norm = [l / max(1, len(c)) for l, c in zip(ll, d["choices"])]
                k += len(d["choices"])
                acc += int(max(range(len(ll)), key=ll.__getitem__) == int(d["gold"]))
                acc_norm += int(max(range(len(norm)), key=norm.__getitem__) == int(d["gold"]))
            metrics.update(acc=acc / len(mc), acc_norm=acc_norm / len(mc))
        if ppl:
            lls = lm.loglikelihood_rolling([(d["text"],) for d in ppl])
            words = sum(len(d["text"].split()) for d in ppl)
            nbytes = sum(len(d["text"].encode("utf-8")) for d in ppl)
            total = sum(lls)
            metrics.update(word_perplexity=math.exp(-total / max(1, words)),
                           byte_perplexity=math.exp(-total / max(1, nbytes)),
                           bits_per_byte=-total / max(1, nbytes) / math.log(2))
        if gen:
            outs = lm.generate_until([(d["query"], {"until": d.get("until", ["\n"]),
                                                    "max_gen_toks": d.get("max_gen_toks", 64)}) for d in gen])
            metrics["exact_match"] = sum(o.strip() == d["answer"].strip() for o, d in zip(outs, gen)) / len(gen)
        return metrics


@torch.inference_mode()
def run_eval_harness(model, tokenizer, model_name, eval_tasks: List[str] = ("hellaswag",), batch_size_per_gpu: int = 1,
                     save_filepath: Optional[Path] = None, limit: Optional[int] = None,
                     bootstrap_iters: int = 100000, dtype=None, cfg=None):
    model.eval()
    if dtype is None:
        dtype = torch.bfloat16 if next(model.parameters()).is_cuda else torch.float32
    model = model.to(dtype)
    harness = EvalHarnessBase(model, tokenizer, model_name, batch_size_per_gpu, cfg)
    results = harness.run_eval(list(eval_tasks), limit, bootstrap_iters)
    if dist.is_main_process():
        if save_filepath is None:
            print(results["results"])
        else:
            print(f"Saving results to {str(save_filepath)!r}")
            with open(save_filepath, "w") as fw:
                fw.write(json.dumps(results))
    return results
