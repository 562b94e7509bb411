"""Every model family builds from its packaged config, runs forward + backward on CPU and produces finite losses and
gradients for all parameters (the reference's per-model tests need GPUs + downloaded corpora; these are the CPU
counterpart on synthetic batches — the multi-rank layout equivalence lives in tests/test_parallel_cpu.py)."""
import pytest
import torch

from libai_b200.config import LazyConfig, instantiate

NLP_OVERRIDES = ["model.cfg.hidden_size=48", "model.cfg.num_attention_heads=4", "model.cfg.hidden_layers=2",
                 "model.cfg.vocab_size=128"]


def _check(model, out_loss):
    assert torch.isfinite(out_loss).all()
    out_loss.backward()
    missing = [n for n, p in model.named_parameters() if p.requires_grad and p.grad is None
               and getattr(p, "main_grad", None) is None]
    assert not missing, missing


def _load(cfg_file, overrides):
    cfg = LazyConfig.load(cfg_file)
    return LazyConfig.apply_overrides(cfg, overrides)


def test_gpt():
    cfg = _load("configs/gpt2_pretrain.py", NLP_OVERRIDES + ["model.cfg.ffn_hidden_size=96", "model.cfg.max_seq_length=32"])
    model = instantiate(cfg.model).train()
    ids = torch.randint(0, 128, (2, 32))
    out = model(ids, ids)
    assert set(out) == {"lm_loss"}
    _check(model, out["lm_loss"])
    model.eval()
    assert model(ids)["prediction_scores"].shape == (2, 32, 128)


def test_bert_pretraining_and_classification():
    cfg = _load("configs/bert_large_pretrain.py", NLP_OVERRIDES + ["model.cfg.intermediate_size=96", "model.cfg.max_position_embeddings=32"])
    model = instantiate(cfg.model).train()
    ids = torch.randint(0, 128, (2, 32))
    mask = torch.ones(2, 32, dtype=torch.long)
    mask[1, 20:] = 0
    out = model(ids, mask, torch.zeros_like(ids), ns_labels=torch.tensor([0, 1]), lm_labels=ids, loss_mask=mask)
    assert {"lm_loss", "sop_loss"} <= set(out)
    _check(model, out["lm_loss"] + out["sop_loss"])
    cfg = _load("configs/bert_classification.py", NLP_OVERRIDES + ["model.cfg.intermediate_size=96", "model.cfg.max_position_embeddings=32"])
    clf = instantiate(cfg.model).train()
    out = clf(ids, mask, torch.zeros_like(ids), labels=torch.tensor([0, 1]))
    _check(clf, out["loss"] if "loss" in out else next(iter(out.values())))


def test_roberta():
    cfg = _load("configs/roberta_pretrain.py", NLP_OVERRIDES + ["model.cfg.intermediate_size=96", "model.cfg.max_position_embeddings=40"])
    model = instantiate(cfg.model).train()
    ids = torch.randint(2, 128, (2, 32))
    mask = torch.ones(2, 32, dtype=torch.long)
    out = model(ids, mask, torch.zeros_like(ids), lm_labels=ids, loss_mask=mask)
    _check(model, out["lm_loss"])


def test_t5():
    cfg = _load("configs/t5_large_pretrain.py", NLP_OVERRIDES + ["model.cfg.intermediate_size=96", "model.cfg.max_position_embeddings=32"])
    model = instantiate(cfg.model).train()
    enc, dec = torch.randint(0, 128, (2, 16)), torch.randint(0, 128, (2, 8))
    out = model(enc, dec, torch.ones(2, 16, 16, dtype=torch.bool), torch.ones(2, 8, 8, dtype=torch.bool).tril(),
                torch.ones(2, 8, 16, dtype=torch.bool), lm_labels=dec, loss_mask=torch.ones(2, 8))
    _check(model, out["masked_lm_loss"])


def test_llama():
    cfg = LazyConfig.load("configs/common/models/llama.py")
    cfg = LazyConfig.apply_overrides(cfg, ["cfg.hidden_layers=2", "cfg.hidden_size=64", "cfg.intermediate_size=128",
                                           "cfg.num_attention_heads=4", "cfg.vocab_size=128", "cfg.max_position_embeddings=32"])
    cfg.model.cfg = cfg.cfg
    model = instantiate(cfg.model).train()
    ids = torch.randint(1, 128, (2, 16))
    _check(model, model(ids, labels=ids)["lm_loss"])


@pytest.mark.parametrize("recipe,overrides,size", [
    ("configs/vit_imagenet.py", ["model.cfg.embed_dim=48", "model.cfg.depth=2", "model.cfg.num_heads=4", "model.cfg.img_size=32", "model.cfg.patch_size=8"], 32),
    ("configs/swin_imagenet.py", ["model.cfg.embed_dim=16", "model.cfg.depths=[2,2]", "model.cfg.num_heads=[2,4]", "model.cfg.img_size=32", "model.cfg.patch_size=2", "model.cfg.window_size=4"], 32),
    ("configs/swinv2_imagenet.py", ["model.cfg.embed_dim=16", "model.cfg.depths=[2,2]", "model.cfg.num_heads=[2,4]", "model.cfg.img_size=32", "model.cfg.patch_size=2", "model.cfg.window_size=4", "model.cfg.pretrained_window_sizes=[0,0]"], 32),
    ("configs/resmlp_imagenet.py", ["model.cfg.embed_dim=48", "model.cfg.depth=2", "model.cfg.img_size=32", "model.cfg.patch_size=8"], 32),
])
def test_vision_models_with_mixup(recipe, overrides, size):
    cfg = _load(recipe, overrides + ["model.cfg.num_classes=10", "dataloader.train.mixup_func.num_classes=10"])
    model = instantiate(cfg.model).train()
    mixup = instantiate(cfg.dataloader.train.mixup_func)
    images, soft = mixup(torch.randn(4, 3, size, size), torch.tensor([1, 2, 3, 4]))
    out = model(images, soft)
    _check(model, out["losses"])
    model.eval()
    assert model(images)["prediction_scores"].shape == (4, 10)
