"""Generation under tensor / pipeline parallelism (reference tests/inference/test_text_generation.py:41-90: TP4, PP4 and
TP2×PP2 on 4 devices, ``use_cache`` on == off).  Here on 4 gloo CPU processes, and additionally every layout must
produce exactly the tokens of the single-process run (parameter initialisation is layout independent)."""
import numpy as np
import pytest

from libai_b200.config import LazyConfig

OVERRIDES = [
    "model.cfg.vocab_size=160", "model.cfg.hidden_size=32", "model.cfg.hidden_layers=4",
    "model.cfg.num_attention_heads=4", "model.cfg.head_size=8", "model.cfg.intermediate_size=64",
    "model.cfg.max_length=10",
]
TEXTS = ["summarize: she is a student", "the quick brown fox", "lazy dog jumps over the tall student she loves"]


def _make_tokenizer(tmp):
    import sentencepiece as spm

    corpus = tmp + "/corpus.txt"
    words = "she is a student tall loves study summarize the quick brown fox jumps over lazy dog".split()
    rng = np.random.default_rng(0)
    with open(corpus, "w") as f:
        f.write("\n".join(" ".join(rng.choice(words, 8)) for _ in range(400)) + "\n")
    spm.SentencePieceTrainer.train(input=corpus, model_prefix=tmp + "/spiece", vocab_size=48, model_type="unigram",
                                   pad_id=0, eos_id=1, unk_id=2, bos_id=-1, hard_vocab_limit=False, minloglevel=2)
    return tmp + "/spiece.model"


def _generate(rank, world, spiece, tp, pp):
    from libai_b200.inference.text_generation import TextGenerationPipeline

    cfg = LazyConfig.load("projects/MT5/configs/t5_inference.py")
    cfg = LazyConfig.apply_overrides(cfg, OVERRIDES + [f"tokenization.tokenizer.vocab_file={spiece}"])
    # 4 encoder + 4 decoder layers are spread over the stages like the reference does (2 x hidden_layers indices)
    pipe = TextGenerationPipeline(cfg, data_parallel=1, tensor_parallel=tp, pipeline_parallel=pp,
                                  pipeline_num_layers=8, mode="random", device="cpu")
    out = {}
    for use_cache in (False, True):
        texts = []
        for t in TEXTS:
            res = pipe(t, use_cache=use_cache, max_length=10)      # postprocessed on the main process only (reference too)
            texts.append(res[0]["generated_text"] if res else None)
        out[use_cache] = texts
    return out


@pytest.mark.parametrize("tp,pp", [(4, 1), (1, 4), (2, 2)])
def test_generation_consistent_across_layouts(tmp_path, tp, pp):
    pytest.importorskip("sentencepiece")
    from tests.dist_utils import run_distributed

    spiece = _make_tokenizer(str(tmp_path))
    single = _generate(0, 1, spiece, 1, 1)
    assert single[False] == single[True]
    res = run_distributed(_generate, 4, spiece, tp, pp, timeout=600)
    r = res[0]                          # the main process holds the decoded text
    assert r[False] == r[True], (tp, pp, r)
    assert r[True] == single[True], (tp, pp, r, single)
