"""HuggingFace → libai_b200 weight conversion, checked numerically against ``transformers`` on random-init models
(model: reference tests/model_loader/*, which download real checkpoints)."""
import pytest
import torch

from libai_b200.config import DictConfig

transformers = pytest.importorskip("transformers")


def _save(hf, tmp_path, name):
    d = str(tmp_path / name)
    hf.save_pretrained(d)
    return d


def test_gpt2_loader(tmp_path):
    from libai_b200.models import GPTForPreTraining
    from libai_b200.models.utils.model_loader import GPT2LoaderHuggerFace, GPT2LoaderLiBai

    torch.manual_seed(0)
    hf = transformers.GPT2LMHeadModel(transformers.GPT2Config(
        n_layer=2, n_embd=64, n_head=4, n_positions=32, vocab_size=96, resid_pdrop=0, embd_pdrop=0, attn_pdrop=0)).eval()
    cfg = DictConfig(dict(hidden_layers=1, vocab_size=1, hidden_size=8, ffn_hidden_size=8, num_attention_heads=1,
                          max_seq_length=8, embedding_dropout_prob=0.1, attention_dropout_prob=0.1, output_dropout_prob=0.1,
                          layernorm_epsilon=1e-5, initializer_range=0.02, use_scaled_init_for_output_weights=True,
                          bias_gelu_fusion=True, bias_dropout_fusion=True, scale_mask_softmax_fusion=True,
                          apply_query_key_layer_scaling=False, apply_residual_post_layernorm=False, amp_enabled=False))
    model = GPT2LoaderHuggerFace(GPTForPreTraining, cfg, _save(hf, tmp_path, "gpt2")).load().eval()
    assert cfg.hidden_layers == 2 and cfg.hidden_size == 64 and cfg.ffn_hidden_size == 256
    ids = torch.randint(0, 96, (2, 16))
    with torch.no_grad():
        assert (model(ids)["prediction_scores"] - hf(ids).logits).abs().max() < 1e-3
    # LiBai-format round trip through the Checkpointer layout
    from libai_b200.utils.checkpoint import Checkpointer

    Checkpointer(model, str(tmp_path / "ckpt")).save("model_0000001")
    again = GPT2LoaderLiBai(GPTForPreTraining, cfg, str(tmp_path / "ckpt")).load().eval()
    with torch.no_grad():
        assert torch.equal(again(ids)["prediction_scores"], model(ids)["prediction_scores"])


def test_bert_loader(tmp_path):
    from libai_b200.models import BertForPreTraining
    from libai_b200.models.utils.model_loader import BertLoaderHuggerFace

    torch.manual_seed(0)
    hf = transformers.BertForPreTraining(transformers.BertConfig(
        vocab_size=96, hidden_size=64, num_hidden_layers=2, num_attention_heads=4, intermediate_size=128,
        max_position_embeddings=32, hidden_dropout_prob=0, attention_probs_dropout_prob=0)).eval()
    cfg = DictConfig(dict(vocab_size=1, hidden_size=8, hidden_layers=1, num_attention_heads=1, intermediate_size=8,
                          hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1, max_position_embeddings=8,
                          num_tokentypes=2, add_pooling_layer=True, initializer_range=0.02, layernorm_eps=1e-5,
                          bias_gelu_fusion=True, bias_dropout_fusion=True, scale_mask_softmax_fusion=True,
                          apply_query_key_layer_scaling=False, apply_residual_post_layernorm=False, amp_enabled=False,
                          add_binary_head=True))
    model = BertLoaderHuggerFace(BertForPreTraining, cfg, _save(hf, tmp_path, "bert")).load().eval()
    assert cfg.apply_residual_post_layernorm is True
    ids = torch.randint(0, 96, (2, 16))
    am = torch.ones(2, 16, dtype=torch.long)
    am[1, 12:] = 0
    tt = torch.zeros(2, 16, dtype=torch.long)
    with torch.no_grad():
        a, b = model(ids, am, tt), hf(input_ids=ids, attention_mask=am, token_type_ids=tt)
    assert (a["prediction_scores"] - b.prediction_logits).abs().max() < 1e-2
    assert (a["seq_relationship_score"] - b.seq_relationship_logits).abs().max() < 1e-3


@pytest.mark.parametrize("family", ["vit", "swin", "swinv2"])
def test_vision_loaders(tmp_path, family):
    from libai_b200.models import SwinTransformer, SwinTransformerV2, VisionTransformer
    from libai_b200.models.utils.model_loader import SwinLoaderHuggerFace, SwinV2LoaderHuggerFace, ViTLoaderHuggerFace

    torch.manual_seed(0)
    x = torch.randn(2, 3, 32, 32)
    if family == "vit":
        hf = transformers.ViTForImageClassification(transformers.ViTConfig(
            image_size=32, patch_size=8, hidden_size=64, num_hidden_layers=2, num_attention_heads=4, intermediate_size=256,
            num_labels=5, hidden_dropout_prob=0, attention_probs_dropout_prob=0)).eval()
        cfg = DictConfig(dict(img_size=224, patch_size=16, in_chans=3, embed_dim=8, depth=1, num_heads=1, mlp_ratio=4.0,
                              drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.0, num_classes=5, loss_func=None))
        model = ViTLoaderHuggerFace(VisionTransformer, cfg, _save(hf, tmp_path, "vit")).load().eval()
    elif family == "swin":
        hf = transformers.SwinForImageClassification(transformers.SwinConfig(
            image_size=32, patch_size=2, embed_dim=16, depths=[2, 2], num_heads=[2, 4], window_size=4, num_labels=5,
            drop_path_rate=0.0)).eval()
        cfg = DictConfig(dict(img_size=224, patch_size=4, in_chans=3, num_classes=5, embed_dim=8, depths=[1], num_heads=[1],
                              window_size=7, mlp_ratio=4.0, qkv_bias=True, qk_scale=None, drop_rate=0.0, drop_path_rate=0.0,
                              ape=False, patch_norm=True, loss_func=None))
        model = SwinLoaderHuggerFace(SwinTransformer, cfg, _save(hf, tmp_path, "swin")).load().eval()
    else:
        hf = transformers.Swinv2ForImageClassification(transformers.Swinv2Config(
            image_size=32, patch_size=2, embed_dim=16, depths=[2, 2], num_heads=[2, 4], window_size=4, num_labels=5,
            drop_path_rate=0.0)).eval()
        cfg = DictConfig(dict(img_size=224, patch_size=4, in_chans=3, num_classes=5, embed_dim=8, depths=[1], num_heads=[1],
                              window_size=7, mlp_ratio=4.0, qkv_bias=True, drop_rate=0.0, drop_path_rate=0.0, ape=False,
                              patch_norm=True, pretrained_window_sizes=[0, 0], loss_func=None))
        model = SwinV2LoaderHuggerFace(SwinTransformerV2, cfg, _save(hf, tmp_path, "swinv2")).load().eval()
    with torch.no_grad():
        assert (model(x)["prediction_scores"] - hf(x).logits).abs().max() < 5e-3


def test_llama_loader_and_generation_parity(tmp_path):
    from libai_b200.models import LlamaForCausalLM
    from libai_b200.models.utils.model_loader import LlamaLoaderHuggerFace

    torch.manual_seed(0)
    hf = transformers.LlamaForCausalLM(transformers.LlamaConfig(
        vocab_size=96, hidden_size=64, intermediate_size=176, num_hidden_layers=2, num_attention_heads=4,
        max_position_embeddings=64, rms_norm_eps=1e-5, bos_token_id=1, eos_token_id=2, pad_token_id=0)).eval()
    cfg = DictConfig(dict(hidden_layers=1, vocab_size=1, hidden_size=8, intermediate_size=8, num_attention_heads=1,
                          max_position_embeddings=8, rms_norm_eps=1e-6, initializer_range=0.02,
                          use_scaled_init_for_output_weights=True, scale_mask_softmax_fusion=False, amp_enabled=False))
    model = LlamaLoaderHuggerFace(LlamaForCausalLM, cfg, _save(hf, tmp_path, "llama")).load().eval()
    ids = torch.randint(3, 96, (2, 8))
    with torch.no_grad():
        assert (model(ids)["logits"] - hf(ids).logits).abs().max() < 1e-4
    ours = model.generate(ids, max_length=16, eos_token_id=None, pad_token_id=0)
    theirs = hf.generate(ids, max_length=16, do_sample=False, eos_token_id=None, pad_token_id=0)
    assert torch.equal(ours, theirs)


def test_roberta_loader(tmp_path):
    """HF RobertaForMaskedLM (random init) → RobertaForPreTraining: masked-LM logits agree, incl. a padded sample
    (RoBERTa's position ids start after `pad_token_id`)."""
    from libai_b200.models import RobertaForPreTraining
    from libai_b200.models.utils.model_loader import RobertaLoaderHuggerFace

    torch.manual_seed(0)
    hf = transformers.RobertaForMaskedLM(transformers.RobertaConfig(
        vocab_size=96, hidden_size=64, num_hidden_layers=2, num_attention_heads=4, intermediate_size=128,
        max_position_embeddings=40, type_vocab_size=1, pad_token_id=1, hidden_dropout_prob=0,
        attention_probs_dropout_prob=0)).eval()
    cfg = DictConfig(dict(vocab_size=1, hidden_size=8, hidden_layers=1, num_attention_heads=1, intermediate_size=8,
                          hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1, max_position_embeddings=8,
                          num_tokentypes=1, add_pooling_layer=False, initializer_range=0.02, layernorm_eps=1e-5,
                          pad_token_id=1, bias_gelu_fusion=True, bias_dropout_fusion=True, scale_mask_softmax_fusion=True,
                          apply_query_key_layer_scaling=False, apply_residual_post_layernorm=False, amp_enabled=False))
    model = RobertaLoaderHuggerFace(RobertaForPreTraining, cfg, _save(hf, tmp_path, "roberta")).load().eval()
    ids = torch.randint(2, 96, (2, 16))
    am = torch.ones(2, 16, dtype=torch.long)
    am[1, 11:] = 0
    ids[1, 11:] = 1                      # padded tail
    with torch.no_grad():
        a = model(ids, am)
        b = hf(input_ids=ids, attention_mask=am)
    valid = am.bool()
    assert (a["prediction_scores"][valid] - b.logits[valid]).abs().max() < 1e-2
