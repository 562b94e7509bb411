"""Eval_LLM harness: request types against a brute-force fp32 reference on a tiny random Llama."""
import json
import math

import torch

from libai_b200.config import DictConfig, LazyConfig
from libai_b200.models.llama_model import LlamaForCausalLM
from libai_b200.utils import distributed as dist
from projects.Eval_LLM.eval_harness import (
    EvalHarnessBase, LocalTask, get_rolling_token_windows, make_disjoint_window,
)


class CharTok:
    eos_token_id = 1
    pad_token_id = 0

    def encode(self, s, add_special_tokens=False):
        return [2 + (ord(c) % 60) for c in s]

    def decode(self, ids):
        return "".join(chr(97 + (int(i) - 2) % 26) for i in ids if int(i) >= 2)


def _tiny():
    dist.setup_dist_util(DictConfig(dict(data_parallel_size=1, tensor_parallel_size=1, pipeline_parallel_size=1,
                                         device_type="cpu")))
    lc = LazyConfig.load("configs/common/models/llama.py")
    lc = LazyConfig.apply_overrides(lc, ["cfg.hidden_layers=2", "cfg.hidden_size=32", "cfg.intermediate_size=64",
                                         "cfg.num_attention_heads=4", "cfg.vocab_size=64",
                                         "cfg.max_position_embeddings=24", "cfg.eos_token_id=1", "cfg.bos_token_id=1",
                                         "cfg.pad_token_id=0", "cfg.max_length=8"])
    cfg = lc.cfg
    torch.manual_seed(0)
    return LlamaForCausalLM(cfg).eval(), cfg


def test_rolling_windows_cover_each_token_once():
    toks = list(range(10, 33))
    wins = [make_disjoint_window(w) for w in get_rolling_token_windows(toks, 1, 8, 1)]
    pred = [t for _, p in wins for t in p]
    assert pred == toks
    for ctx, p in wins:
        assert len(ctx) >= 1 and len(ctx) + len(p) - 1 <= 8


def test_loglikelihood_matches_bruteforce():
    model, cfg = _tiny()
    tok = CharTok()
    lm = EvalHarnessBase(model, tok, "tiny", batch_size=3, cfg=cfg)
    reqs = [("hello wor", "ld"), ("", "abc"), ("the quick brown fox jumps over", " the lazy dog"), ("a", "b")]
    got = lm.loglikelihood(reqs)
    for (c, k), (score, greedy) in zip(reqs, got):
        ce = [tok.eos_token_id] if c == "" else tok.encode(c)
        ke = tok.encode(k)
        full = (ce + ke)[-(cfg.max_position_embeddings + 1):]
        with torch.no_grad():
            logits = model(torch.tensor([full[:-1]]))["logits"][0].float()
        logp = torch.log_softmax(logits, -1)
        want = sum(logp[len(full) - 1 - len(ke) + i, t].item() for i, t in enumerate(ke))
        assert abs(want - score) < 1e-3, (c, k, want, score)
        assert isinstance(greedy, bool)


def test_rolling_and_local_task(tmp_path):
    model, cfg = _tiny()
    lm = EvalHarnessBase(model, CharTok(), "tiny", batch_size=2, cfg=cfg)
    text = "lorem ipsum dolor sit amet consectetur adipiscing elit sed do"
    (ll,) = lm.loglikelihood_rolling([(text,)])
    assert ll < 0 and math.isfinite(ll)
    # uniform-ish random model: per-token nll close to log(vocab)
    assert abs(-ll / len(text) - math.log(cfg.vocab_size)) < 0.5

    p = tmp_path / "toy.jsonl"
    with open(p, "w") as f:
        f.write(json.dumps({"query": "two plus two is", "choices": [" four", " five", " six"], "gold": 0}) + "\n")
        f.write(json.dumps({"text": text}) + "\n")
        f.write(json.dumps({"query": "say", "until": ["z"], "answer": "x", "max_gen_toks": 4}) + "\n")
    res = lm.run_eval([str(p)], limit=None)
    m = res["results"]["toy"]
    assert set(m) >= {"acc", "acc_norm", "word_perplexity", "bits_per_byte", "exact_match"}
    assert res["config"]["model"] == "tiny"
    assert LocalTask(p).name == "toy"
