"""Shared driver of the dist_infer_* scripts: build/load a HF causal LM, tensor-parallelise it, generate."""
import argparse
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))

import torch  # noqa: E402

from libai_b200.utils import distributed as dist  # noqa: E402
from projects.mock_transformers import init_env  # noqa: E402

TINY = {
    "gpt2": dict(n_embd=64, n_layer=2, n_head=4, vocab_size=128, n_positions=64),
    "llama": dict(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
                  num_key_value_heads=2, vocab_size=128, max_position_embeddings=64),
    "opt": dict(hidden_size=64, ffn_dim=128, num_hidden_layers=2, num_attention_heads=4, vocab_size=128,
                max_position_embeddings=64, word_embed_proj_dim=64),
    "bloom": dict(hidden_size=64, n_layer=2, n_head=4, vocab_size=128),
}


def tiny_model(model_type, seed=0):
    """A randomly initialised two-layer model of the family (offline smoke runs and the unit tests)."""
    from transformers import AutoConfig, AutoModelForCausalLM

    torch.manual_seed(seed)
    config = AutoConfig.for_model(model_type, **TINY[model_type])
    return AutoModelForCausalLM.from_config(config).eval()


def run(model_type, default_model, default_prompt, argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default=default_model, help="HF hub name or local path")
    ap.add_argument("--prompt", default=default_prompt)
    ap.add_argument("--max_length", type=int, default=30)
    ap.add_argument("--tensor_parallel", type=int, default=None, help="default: all launched processes")
    ap.add_argument("--random", action="store_true", help="tiny random-weight model, token-id prompt (no downloads)")
    args = ap.parse_args(argv)

    if "RANK" in os.environ and not torch.distributed.is_initialized():
        torch.distributed.init_process_group("nccl" if torch.cuda.is_available() else "gloo")
    if torch.cuda.is_available():
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    init_env.setup(args.tensor_parallel)

    if args.random:
        model = init_env.load_parallel(tiny_model(model_type), model_type)
        input_ids = torch.arange(3, 11)[None].to(next(model.parameters()).device)
        out = model.generate(input_ids, max_length=args.max_length, do_sample=False, pad_token_id=0)
        if dist.is_main_process():
            print(out.tolist())
        return out
    from projects.mock_transformers.mock_tokenization import wrap_tokenizer
    from transformers import AutoTokenizer

    model = init_env.load_parallel(args.model, model_type, trust_remote_code=True)
    tokenizer = wrap_tokenizer(AutoTokenizer.from_pretrained(args.model, use_fast=False, trust_remote_code=True))
    batch = tokenizer(args.prompt, return_tensors="of", is_global=True)
    generated = model.generate(batch["input_ids"], max_length=args.max_length, do_sample=False)
    text = tokenizer.batch_decode(generated, skip_special_tokens=True)
    if dist.is_main_process():
        print(text)
    return text
