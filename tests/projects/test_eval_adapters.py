"""projects/{Llama,ChatGLM}/utils/eval_adapter.py: checkpoint → harness → local JSONL task, end to end on CPU."""
import json

import numpy as np
import pytest
import torch

from libai_b200.config import DictConfig
from libai_b200.utils import distributed as dist


def _spm_model(tmp_path, vocab=64):
    spm = pytest.importorskip("sentencepiece")
    words = "two plus four five six is the answer lorem ipsum dolor sit amet say hello world".split()
    rng = np.random.default_rng(0)
    corpus = tmp_path / "corpus.txt"
    corpus.write_text("\n".join(" ".join(rng.choice(words, 8)) for _ in range(400)) + "\n")
    spm.SentencePieceTrainer.train(input=str(corpus), model_prefix=str(tmp_path / "tok"), vocab_size=vocab,
                                   model_type="unigram", unk_id=0, bos_id=1, eos_id=2, pad_id=-1,
                                   hard_vocab_limit=False, minloglevel=2)
    return str(tmp_path / "tok.model")


def test_llama_eval_adapter(tmp_path):
    from libai_b200.utils.checkpoint import Checkpointer
    from projects.Llama.configs.llama_config import cfg
    from projects.Llama.llama import LlamaForCausalLM
    from projects.Llama.utils import eval_adapter

    saved = dict(cfg)
    try:
        cfg.update(hidden_size=32, intermediate_size=64, num_attention_heads=4, hidden_layers=2, vocab_size=64,
                   max_position_embeddings=64, amp_enabled=False, max_length=8)
        dist.reset_dist_util()
        dist.setup_dist_util(DictConfig(dict(data_parallel_size=1, tensor_parallel_size=1, pipeline_parallel_size=1,
                                             device_type="cpu")))
        torch.manual_seed(0)
        model = LlamaForCausalLM(cfg)
        Checkpointer(model, str(tmp_path / "ckpt")).save("model_final")
        task = tmp_path / "toy.jsonl"
        with open(task, "w") as f:
            f.write(json.dumps({"query": "two plus two is", "choices": [" four", " five", " six"], "gold": 0}) + "\n")
            f.write(json.dumps({"text": "lorem ipsum dolor sit amet"}) + "\n")
        dist.reset_dist_util()
        out = tmp_path / "res.json"
        res = eval_adapter.main(["--model-path", str(tmp_path / "ckpt"), "--format", "libai", "--tokenizer-path",
                                 _spm_model(tmp_path), "--tasks", str(task), "--device", "cpu", "--save", str(out)])
        m = res["results"]["toy"]
        assert 0.0 <= m["acc"] <= 1.0 and m["word_perplexity"] > 1.0
        assert json.loads(out.read_text())["results"]["toy"]["acc"] == m["acc"]
    finally:
        cfg.clear()
        cfg.update(saved)
        dist.reset_dist_util()


def test_chatglm_eval_adapter_binds_project_classes():
    from projects.ChatGLM.chatglm import ChatGLMForConditionalGeneration
    from projects.ChatGLM.utils import eval_adapter
    from projects.ChatGLM.utils.chatglm_loader import ChatGLMLoaderHuggerFace, ChatGLMLoaderLiBai

    assert eval_adapter.ChatGLMForConditionalGeneration is ChatGLMForConditionalGeneration
    assert eval_adapter.ChatGLMLoaderHuggerFace is ChatGLMLoaderHuggerFace and eval_adapter.ChatGLMLoaderLiBai is ChatGLMLoaderLiBai
    with pytest.raises(SystemExit):
        eval_adapter.main(["--format", "libai"])            # --model-path is required
