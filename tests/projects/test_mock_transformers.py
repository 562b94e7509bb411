"""mock_transformers: a tensor-parallelised HF model reproduces the single-process logits and generation."""
import numpy as np
import pytest
import torch

from tests.dist_utils import run_distributed


def _tp_worker(rank, world, model_type):
    from projects.mock_transformers import init_env
    from projects.mock_transformers._common import tiny_model

    init_env.setup(world, "cpu")
    ids = torch.arange(3, 19).view(2, 8)
    ref = tiny_model(model_type)
    with torch.no_grad():
        want = ref(ids).logits
        want_gen = ref.generate(ids[:1], max_length=14, do_sample=False, pad_token_id=0)
    model = init_env.load_parallel(tiny_model(model_type), model_type, dtype=torch.float32, device="cpu")
    assert len(model._tp_replaced) >= 8
    with torch.no_grad():
        got = model(ids).logits
        got_gen = model.generate(ids[:1], max_length=14, do_sample=False, pad_token_id=0)
    return {"err": float((got - want).abs().max()), "gen": bool((got_gen == want_gen).all()),
            "n": len(model._tp_replaced)}


@pytest.mark.parametrize("model_type", ["gpt2", "llama", "opt", "bloom"])
def test_tp2_matches_single(model_type):
    out = run_distributed(_tp_worker, 2, model_type)
    for r in out:
        assert r["err"] < 2e-4, (model_type, r)
        assert r["gen"], (model_type, r)
