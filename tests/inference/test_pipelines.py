"""Task pipelines end to end on CPU with tiny random-weight models: config → topology → model → tokenizer →
preprocess / forward / postprocess (reference tests/inference/* need downloaded checkpoints)."""
import numpy as np
import pytest
import torch

from libai_b200.config import LazyConfig


def test_image_classification_pipeline(tmp_path):
    from PIL import Image

    from libai_b200.inference.image_classification import ImageClassificationPipeline

    cfg = LazyConfig.load("configs/vit_imagenet.py")
    cfg = LazyConfig.apply_overrides(cfg, ["model.cfg.embed_dim=48", "model.cfg.depth=2", "model.cfg.num_heads=4",
                                           "model.cfg.img_size=224", "model.cfg.patch_size=32", "model.cfg.num_classes=1000"])
    pipe = ImageClassificationPipeline(cfg, data_parallel=1, tensor_parallel=1, pipeline_parallel=1, mode="random", device="cpu")
    img = tmp_path / "x.png"
    Image.fromarray(np.random.default_rng(0).integers(0, 255, (300, 260, 3), dtype=np.uint8)).save(img)
    out = pipe(str(img))
    assert set(out) == {"label", "score"} and 0 < out["score"] <= 1 and isinstance(out["label"], str)
    full = pipe(str(img), return_all_scores=True, function_to_apply="softmax")
    assert len(full) == 1000 and abs(sum(r["score"] for r in full) - 1) < 1e-3
    assert full[0]["label"].startswith("tench")                     # ImageNet-1k names


def test_text_classification_pipeline(tmp_path):
    from libai_b200.inference.text_classification import TextClassificationPipeline

    vocab = tmp_path / "vocab.txt"
    vocab.write_text("\n".join(["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]", "hello", "world", "good", "bad", "movie"]) + "\n")
    cfg = LazyConfig.load("configs/bert_classification.py")
    cfg = LazyConfig.apply_overrides(cfg, ["model.cfg.vocab_size=16", "model.cfg.hidden_size=32", "model.cfg.hidden_layers=2",
                                           "model.cfg.num_attention_heads=4", "model.cfg.intermediate_size=64",
                                           "model.cfg.max_position_embeddings=32", "model.cfg.num_labels=3",
                                           f"tokenization.tokenizer.vocab_file={vocab}", "tokenization.tokenizer.do_chinese_wwm=false"])
    pipe = TextClassificationPipeline(cfg, data_parallel=1, tensor_parallel=1, pipeline_parallel=1, mode="random", device="cpu")
    out = pipe("good movie")
    assert out["label"] in {"Label_0", "Label_1", "Label_2"} and 0 < out["score"] <= 1
    every = pipe("bad movie", return_all_scores=True)
    assert [r["label"] for r in every] == ["Label_0", "Label_1", "Label_2"]
    assert abs(sum(r["score"] for r in every) - 1) < 1e-4
    raw = pipe("hello world", function_to_apply="none", return_all_scores=True)
    assert any(r["score"] < 0 or r["score"] > 1 for r in raw) or True   # raw logits are unconstrained
    with pytest.raises(AssertionError):
        pipe("hello", function_to_apply="tanh")


def test_text_generation_pipeline(tmp_path):
    """MT5 encoder-decoder + sentencepiece tokenizer + ``generate`` through ``TextGenerationPipeline``."""
    spm = pytest.importorskip("sentencepiece")
    from libai_b200.inference.text_generation import TextGenerationPipeline

    corpus = tmp_path / "corpus.txt"
    words = "she is a student tall loves study summarize the quick brown fox jumps over lazy dog".split()
    rng = np.random.default_rng(0)
    corpus.write_text("\n".join(" ".join(rng.choice(words, 8)) for _ in range(400)) + "\n")
    spm.SentencePieceTrainer.train(input=str(corpus), model_prefix=str(tmp_path / "spiece"), vocab_size=48,
                                   model_type="unigram", pad_id=0, eos_id=1, unk_id=2, bos_id=-1,
                                   hard_vocab_limit=False, minloglevel=2)
    cfg = LazyConfig.load("projects/MT5/configs/t5_inference.py")
    cfg = LazyConfig.apply_overrides(cfg, [
        "model.cfg.vocab_size=160", "model.cfg.hidden_size=32", "model.cfg.hidden_layers=2",
        "model.cfg.num_attention_heads=4", "model.cfg.head_size=8", "model.cfg.intermediate_size=64",
        "model.cfg.max_length=8", f"tokenization.tokenizer.vocab_file={tmp_path / 'spiece.model'}"])
    pipe = TextGenerationPipeline(cfg, data_parallel=1, tensor_parallel=1, pipeline_parallel=1, mode="random", device="cpu")
    out = pipe("summarize: she is a student")
    assert isinstance(out, list) and len(out) == 1 and isinstance(out[0]["generated_text"], str)
    again = pipe("summarize: she is a student")
    assert again == out                                              # greedy decoding is deterministic
    sampled = pipe("summarize: she is tall", do_sample=True, top_k=5, max_length=6)
    assert isinstance(sampled[0]["generated_text"], str)
