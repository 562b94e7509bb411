"""One tiny-model check per project: HuggingFace numerical parity where `transformers` has the architecture,
forward/backward + config loading otherwise (the reference's per-project READMEs are the only spec it ships)."""
import os

import pytest
import torch

from libai_b200.config import DictConfig, LazyConfig, instantiate
from libai_b200.utils import distributed as dist

transformers = pytest.importorskip("transformers")


@pytest.fixture(autouse=True)
def _topo():
    dist.setup_dist_util(DictConfig(dict(data_parallel_size=1, tensor_parallel_size=1, pipeline_parallel_size=1,
                                         device_type="cpu")))


def _save(hf, tmp_path, name):
    d = str(tmp_path / name)
    hf.save_pretrained(d)
    return d


GEN = dict(is_encoder_decoder=False, max_length=20, min_length=0, do_sample=False, early_stopping=False, num_beams=1,
           num_beam_groups=1, diversity_penalty=0.0, temperature=1.0, top_k=50, top_p=1.0, typical_p=1.0,
           repetition_penalty=1.0, length_penalty=1.0, no_repeat_ngram_size=0, encoder_no_repeat_ngram_size=0,
           num_return_sequences=1, chunk_size_feed_forward=0, output_scores=False, use_cache=True)


# ------------------------------------------------------------------------------------------------ Llama family
def test_llama_adapter_trains_only_the_adapter():
    from projects.Llama.adapter.adapter_model import LlamaForCausalLM

    cfg = DictConfig(dict(hidden_layers=3, vocab_size=100, hidden_size=64, intermediate_size=128, num_attention_heads=4,
                          max_position_embeddings=64, rms_norm_eps=1e-5, initializer_range=0.02,
                          use_scaled_init_for_output_weights=True, scale_mask_softmax_fusion=False, amp_enabled=False,
                          adapter_len=4, adapter_layer=2, bos_token_id=1, eos_token_id=2, pad_token_id=0))
    m = LlamaForCausalLM(cfg).freeze_backbone()
    trainable = sorted(n for n, p in m.named_parameters() if p.requires_grad)
    assert trainable and all("adapter" in n or "gate" in n for n in trainable)
    ids = torch.randint(3, 100, (2, 8))
    m(ids, labels=ids)["lm_loss"].backward()
    assert m.model.adapter_query.weight.grad is not None
    assert m.generate(ids, max_length=12).shape == (2, 12)
    assert LazyConfig.load("projects/Llama/adapter/adapter_sft.py").model.cfg.adapter_len > 0


@pytest.mark.parametrize("recipe,has_qkv_bias", [
    ("projects/Llama/configs/llama_sft.py", False), ("projects/Aquila/configs/aquila_sft.py", False),
    ("projects/Baichuan/configs/baichuan_sft.py", False), ("projects/Qwen/configs/qwen2_sft.py", True)])
def test_llama_family_sft_recipes(recipe, has_qkv_bias):
    c = LazyConfig.load(recipe)
    c = LazyConfig.apply_overrides(c, ["model.cfg.hidden_layers=2", "model.cfg.hidden_size=64", "model.cfg.intermediate_size=128",
                                       "model.cfg.num_attention_heads=4", "model.cfg.vocab_size=128",
                                       "model.cfg.max_position_embeddings=32"])
    if "num_key_value_heads" in c.model.cfg:
        c.model.cfg.num_key_value_heads = 4
    m = instantiate(c.model)
    ids = torch.randint(1, 128, (2, 8))
    out = m(ids, labels=ids)
    out["lm_loss"].backward()
    assert 3.5 < float(out["lm_loss"].detach()) < 6.5
    assert any("query_key_value.bias" in n for n, _ in m.named_parameters()) == has_qkv_bias


def test_qwen2_and_baichuan_hf_loaders(tmp_path):
    from projects.Baichuan.baichuan import BaichuanForCausalLM
    from projects.Baichuan.utils.baichuan_loader import BaichuanLoaderHuggerFace
    from projects.Qwen.qwen2 import Qwen2ForCausalLM
    from projects.Qwen.utils.qwen2_loader import Qwen2LoaderHuggerFace

    torch.manual_seed(0)
    hf = transformers.Qwen2ForCausalLM(transformers.Qwen2Config(
        vocab_size=96, hidden_size=64, intermediate_size=176, num_hidden_layers=2, num_attention_heads=4,
        num_key_value_heads=4, max_position_embeddings=64, rms_norm_eps=1e-6, tie_word_embeddings=False)).eval()
    base = dict(hidden_layers=1, vocab_size=1, hidden_size=8, intermediate_size=8, num_attention_heads=1,
                max_position_embeddings=8, rms_norm_eps=1e-6, initializer_range=0.02,
                use_scaled_init_for_output_weights=False, scale_mask_softmax_fusion=False, amp_enabled=False)
    m = Qwen2LoaderHuggerFace(Qwen2ForCausalLM, DictConfig(dict(base, qkv_bias=True)), _save(hf, tmp_path, "qwen")).load().eval()
    ids = torch.randint(3, 96, (2, 8))
    with torch.no_grad():
        assert (m(ids)["logits"] - hf(ids).logits).abs().max() < 1e-4

    # Baichuan checkpoints store q/k/v fused as W_pack = [q | k | v]; build one from a Llama model
    llama = transformers.LlamaForCausalLM(transformers.LlamaConfig(
        vocab_size=96, hidden_size=64, intermediate_size=176, num_hidden_layers=2, num_attention_heads=4,
        max_position_embeddings=64, rms_norm_eps=1e-6)).eval()
    d = _save(llama, tmp_path, "baichuan")
    from safetensors.torch import load_file, save_file

    sd = load_file(os.path.join(d, "model.safetensors"))
    for i in range(2):
        p = f"model.layers.{i}.self_attn."
        sd[p + "W_pack.weight"] = torch.cat([sd.pop(p + f"{n}_proj.weight") for n in "qkv"], dim=0)
    save_file(sd, os.path.join(d, "model.safetensors"), metadata={"format": "pt"})
    b = BaichuanLoaderHuggerFace(BaichuanForCausalLM, DictConfig(dict(base)), d).load().eval()
    with torch.no_grad():
        assert (b(ids)["logits"] - llama(ids).logits).abs().max() < 1e-4


# ------------------------------------------------------------------------------------------------ HF parity
def test_mt5_loader_logits_greedy_and_beam(tmp_path):
    from projects.MT5.mt5_model import MT5Model
    from projects.MT5.utils.mt5_loader import T5LoaderHuggerFace

    torch.manual_seed(0)
    for ff, tie in (("relu", True), ("gated-gelu", False)):
        hf = transformers.T5ForConditionalGeneration(transformers.T5Config(
            vocab_size=96, d_model=64, d_kv=16, d_ff=128, num_layers=2, num_heads=4, relative_attention_num_buckets=8,
            dropout_rate=0.0, feed_forward_proj=ff, tie_word_embeddings=tie, scale_decoder_outputs=tie, decoder_start_token_id=0, eos_token_id=1,
            pad_token_id=0)).eval()
        cfg = DictConfig(dict(GEN, vocab_size=1, hidden_size=8, hidden_layers=1, num_attention_heads=1, head_size=8,
                              intermediate_size=8, hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1,
                              embedding_dropout_prob=0.1, relative_attention_num_buckets=4, initializer_range=1.0,
                              layernorm_eps=1e-6, amp_enabled=False, model_type="mt5", eos_token_id=1, padding_idx=0,
                              is_encoder_decoder=True, tie_word_embeddings=True, pad_token_id=0,
                              decoder_start_token_id=0, max_length=12))
        m = T5LoaderHuggerFace(MT5Model, cfg, _save(hf, tmp_path, ff), hidden_dropout_prob=0.0,
                               attention_probs_dropout_prob=0.0, embedding_dropout_prob=0.0).load().eval()
        enc, dec = torch.randint(2, 96, (2, 10)), torch.randint(2, 96, (2, 6))
        am = torch.ones(2, 10, dtype=torch.long)
        am[1, 7:] = 0
        with torch.no_grad():
            assert (m(enc, dec, am)["logits"] - hf(input_ids=enc, attention_mask=am, decoder_input_ids=dec).logits).abs().max() < 2e-4
        ours = m.generate(enc, encoder_attn_mask=am, max_length=10)
        theirs = hf.generate(enc, attention_mask=am, max_length=10, do_sample=False)
        assert torch.equal(ours, theirs[:, :ours.shape[1]])
        ours_b = m.generate(enc, encoder_attn_mask=am, max_length=8, num_beams=3)
        theirs_b = hf.generate(enc, attention_mask=am, max_length=8, num_beams=3, do_sample=False, early_stopping=False)
        assert ours_b.tolist() == theirs_b[:, :ours_b.shape[1]].tolist()


def test_magicprompt_gpt2_cached_generation_matches_hf(tmp_path):
    from libai_b200.models.utils.model_loader import GPT2LoaderHuggerFace
    from projects.MagicPrompt.gpt2 import GPTModel

    torch.manual_seed(0)
    hf = transformers.GPT2LMHeadModel(transformers.GPT2Config(
        n_layer=2, n_embd=64, n_head=4, n_positions=64, vocab_size=96, resid_pdrop=0, embd_pdrop=0, attn_pdrop=0,
        activation_function="gelu")).eval()
    cfg = DictConfig(dict(GEN, hidden_layers=1, vocab_size=1, hidden_size=8, ffn_hidden_size=8, num_attention_heads=1,
                          max_seq_length=8, embedding_dropout_prob=0., attention_dropout_prob=0., output_dropout_prob=0.,
                          layernorm_epsilon=1e-5, initializer_range=0.02, use_scaled_init_for_output_weights=True,
                          bias_gelu_fusion=True, bias_dropout_fusion=True, scale_mask_softmax_fusion=True,
                          apply_query_key_layer_scaling=False, apply_residual_post_layernorm=False, amp_enabled=False,
                          eos_token_id=95, pad_token_id=0, bos_token_id=95))
    m = GPT2LoaderHuggerFace(GPTModel, cfg, _save(hf, tmp_path, "gpt2")).load().eval()
    ids = torch.randint(1, 90, (2, 6))
    ours = m.generate(ids, max_length=16, eos_token_id=None)
    theirs = hf.generate(ids, max_length=16, do_sample=False, eos_token_id=None, pad_token_id=0)
    assert torch.equal(ours, theirs)


def test_bloom_matches_hf(tmp_path):
    from projects.BLOOM.modeling.bloom_model import BloomForCausalLM
    from projects.BLOOM.utils.model_loader import BlooMLoaderHuggerFace

    torch.manual_seed(0)
    hf = transformers.BloomForCausalLM(transformers.BloomConfig(
        vocab_size=96, hidden_size=64, n_layer=2, n_head=8, hidden_dropout=0.0, attention_dropout=0.0, pad_token_id=3,
        bos_token_id=1, eos_token_id=2)).eval()
    cfg = LazyConfig.load("projects/BLOOM/configs/bloom_inference.py").cfg
    m = BlooMLoaderHuggerFace(BloomForCausalLM, cfg, _save(hf, tmp_path, "bloom")).load().eval()
    ids, am = torch.randint(4, 96, (2, 7)), torch.ones(2, 7, dtype=torch.long)
    with torch.no_grad():
        assert (m(ids, am)["logits"] - hf(ids, attention_mask=am).logits).abs().max() < 1e-4
    ours = m.generate(ids, attention_mask=am, max_length=14, eos_token_id=None)
    theirs = hf.generate(ids, attention_mask=am, max_length=14, do_sample=False, eos_token_id=None, pad_token_id=3)
    assert torch.equal(ours, theirs)


def test_convnext_trains_and_matches_hf():
    c = LazyConfig.load("projects/ConvNeXT/configs/convnext_imagenet.py")
    c = LazyConfig.apply_overrides(c, ["model.cfg.hidden_sizes=[16,32,64,128]", "model.cfg.depths=[1,1,2,1]",
                                       "model.cfg.num_labels=10", "dataloader.train.mixup_func.num_classes=10"])
    m = instantiate(c.model).train()
    mix = instantiate(c.dataloader.train.mixup_func)
    x, y = mix(torch.randn(4, 3, 64, 64), torch.tensor([1, 2, 3, 4]))
    m(x, y)["losses"].backward()
    hf = transformers.ConvNextForImageClassification(transformers.ConvNextConfig(
        hidden_sizes=[16, 32, 64, 128], depths=[1, 1, 2, 1], num_labels=10)).eval()
    sd = {}
    for k, v in hf.state_dict().items():
        if "embeddings.layernorm" in k or "layers." in k:
            k = k.replace(".layernorm.weight", ".layernorm.norm.weight").replace(".layernorm.bias", ".layernorm.norm.bias")
        k = k.replace("downsampling_layer.0.weight", "downsampling_layer.0.norm.weight")
        k = k.replace("downsampling_layer.0.bias", "downsampling_layer.0.norm.bias")
        sd[k] = v
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all("loss" in k for k in missing), (missing, unexpected)
    m.eval()
    xx = torch.randn(2, 3, 64, 64)
    with torch.no_grad():
        assert (m(xx)["prediction_scores"] - hf(xx).logits).abs().max() < 1e-4


# ------------------------------------------------------------------------------------------------ GLM / ChatGLM / PaLM
def test_glm_cached_generation_equals_uncached_and_trains():
    from projects.GLM.modeling_glm import GLMForConditionalGeneration
    from projects.GLM.tokenizer.glm_tokenizer import GLMTokenizerMixin

    torch.manual_seed(0)
    cfg = DictConfig(dict(GEN, num_layers=2, vocab_size=128, hidden_size=64, num_attention_heads=4, max_sequence_length=64,
                          embedding_dropout_prob=0.0, attention_dropout_prob=0.0, output_dropout_prob=0.0,
                          layernorm_epsilon=1e-5, initializer_range=0.02, use_scaled_init_for_output_weights=True,
                          bias_gelu_fusion=True, bias_dropout_fusion=True, scale_mask_softmax_fusion=False,
                          apply_query_key_layer_scaling=False, amp_enabled=False, block_position_encoding=True,
                          attention_scale=1.0, padding_idx=None, eos_token_id=5, pad_token_id=0, max_length=30))
    m = GLMForConditionalGeneration(cfg).eval()

    class Host:
        mask_token_id, pad_token_id = 7, 0

        def convert_tokens_to_ids(self, t):
            return {"<|startofpiece|>": 9, "<|endofpiece|>": 5, "[gMASK]": 8, "[sMASK]": 6}[t]

        def encode(self, t):
            return [20, 21]

    class Tok(GLMTokenizerMixin, Host):
        pass

    tok = Tok()
    ctx, am = torch.tensor([[11, 12, 13, 7, 14, 15]]), torch.ones(1, 6, dtype=torch.long)
    inp = tok.build_inputs_for_generation({"input_ids": ctx, "attention_mask": am}, max_gen_length=8)
    kw = dict(position_ids=inp["position_ids"], generation_attention_mask=inp["generation_attention_mask"], max_length=12,
              eos_token_id=None)
    assert torch.equal(m.generate(inp["input_ids"], use_cache=True, **kw), m.generate(inp["input_ids"], use_cache=False, **kw))
    tr = tok.build_inputs_for_generation({"input_ids": ctx, "attention_mask": am}, targets=["x"], max_gen_length=8)
    loss = m.train()(**tr)["lm_loss"]
    loss.backward()
    assert torch.isfinite(loss)


def test_chatglm_full_prefix_and_lora():
    c = LazyConfig.load("projects/ChatGLM/configs/chatglm_config.py")
    tiny = ["cfg.num_layers=2", "cfg.hidden_size=64", "cfg.ffn_hidden_size=96", "cfg.num_attention_heads=4", "cfg.kv_channels=16",
            "cfg.multi_query_group_num=2", "cfg.padded_vocab_size=128", "cfg.seq_length=32"]
    c = LazyConfig.apply_overrides(c, tiny)
    c.model.cfg = c.cfg
    m = instantiate(c.model)
    ids = torch.randint(1, 128, (2, 10))
    out = m(ids, labels=ids)
    loss = out["lm_loss"] if "lm_loss" in out else out["loss"]
    loss.backward()
    assert torch.isfinite(loss)
    m.eval()
    g1 = m.generate(ids[:, :5], max_length=9, eos_token_id=None, use_cache=True)
    g2 = m.generate(ids[:, :5], max_length=9, eos_token_id=None, use_cache=False)
    assert torch.equal(g1, g2)

    from projects.ChatGLM.lora.lora_model import LoraModel

    base = instantiate(c.model)
    lora = LoraModel(base, DictConfig(dict(r=4, lora_alpha=8, lora_dropout=0.0, target_modules=["query_key_value"],
                                          fan_in_fan_out=False, bias="none", modules_to_save=None)), "default")
    trainable = [n for n, p in lora.named_parameters() if p.requires_grad]
    assert trainable and all("lora_" in n for n in trainable)
    out = lora(ids, labels=ids)
    (out["lm_loss"] if "lm_loss" in out else out["loss"]).backward()
    assert lora.lora_state_dict() and all("lora_" in k for k in lora.lora_state_dict())
    # merging the adapters into the base weights must not change the function
    lora.eval()
    with torch.no_grad():
        for n, p in lora.named_parameters():
            if "lora_B" in n:
                p.normal_(std=0.05)
        key = "lm_loss" if "lm_loss" in out else "loss"
        before = lora(ids, labels=ids)[key]
        lora.merge_adapter()
        assert torch.allclose(before, lora(ids, labels=ids)[key], atol=1e-4)


def test_palm():
    c = LazyConfig.load("projects/PaLM/configs/palm_pretrain.py")
    c = LazyConfig.apply_overrides(c, ["model.cfg.dim=64", "model.cfg.depth=2", "model.cfg.dim_head=16", "model.cfg.num_heads=4",
                                       "model.cfg.vocab_size=128"])
    m = instantiate(c.model)
    ids = torch.randint(0, 128, (2, 12))
    out = m(ids, ids)
    out["lm_loss"].backward()
    assert 3.5 < float(out["lm_loss"].detach()) < 6.5
    assert LazyConfig.load("projects/PaLM/configs/models/palm_62b.py").model.cfg.dim == 8192


# ------------------------------------------------------------------------------------------------ vision SSL / text
def test_mae_pretrain_and_finetune():
    c = LazyConfig.load("projects/MAE/configs/mae_pretraining.py")
    c = LazyConfig.apply_overrides(c, ["model.img_size=32", "model.patch_size=8", "model.embed_dim=48", "model.depth=2",
                                       "model.num_heads=4", "model.decoder_embed_dim=32", "model.decoder_depth=1",
                                       "model.decoder_num_heads=4"])
    m = instantiate(c.model).train()
    m(torch.randn(2, 3, 32, 32))["losses"].backward()
    o = m.eval()(torch.randn(2, 3, 32, 32))
    assert o["pred"].shape == (2, 16, 192) and m.unpatchify(o["pred"]).shape == (2, 3, 32, 32)
    assert (o["mask"].sum(1) == 12).all()                      # 75 % of 16 patches masked
    c = LazyConfig.load("projects/MAE/configs/mae_finetune.py")
    c = LazyConfig.apply_overrides(c, ["model.img_size=32", "model.patch_size=8", "model.embed_dim=48", "model.depth=2",
                                       "model.num_heads=4", "model.num_classes=10"])
    v = instantiate(c.model).train()
    from projects.MAE.utils.lr_decay import param_groups_lrd

    scales = sorted({round(g["lr_scale"], 3) for g in param_groups_lrd(v, 0.05, 0.65)})
    assert scales[-1] == 1.0 and len(scales) == 4              # embeddings, 2 blocks, head
    out = v(torch.randn(2, 3, 32, 32), torch.nn.functional.one_hot(torch.tensor([1, 2]), 10).float())
    assert torch.isfinite(out["losses"])


def test_mocov3():
    c = LazyConfig.load("projects/MOCOV3/configs/moco_pretrain.py")
    over = []
    for enc in ("base_encoder", "momentum_encoder"):
        over += [f"model.{enc}.img_size=32", f"model.{enc}.patch_size=8", f"model.{enc}.embed_dim=48", f"model.{enc}.depth=2",
                 f"model.{enc}.num_heads=4"]
    c = LazyConfig.apply_overrides(c, over + ["model.dim=16", "model.mlp_dim=32"])
    m = instantiate(c.model).train()
    out = m(torch.randn(4, 6, 32, 32))
    out["losses"].backward()
    assert torch.isfinite(out["losses"]) and int(m.cu_iter) == 1
    assert not m.base_encoder.pos_embed.requires_grad            # fixed sin-cos positions
    assert all(p.grad is None for p in m.momentum_encoder.parameters())


def test_simcse_unsup():
    from projects.SimCSE.modeling.simcse_unsup import Simcse_unsup

    c = LazyConfig.load("projects/SimCSE/config/config_simcse_unsup.py")
    c = LazyConfig.apply_overrides(c, ["model.cfg.hidden_size=48", "model.cfg.hidden_layers=2", "model.cfg.num_attention_heads=4",
                                       "model.cfg.intermediate_size=96", "model.cfg.vocab_size=128",
                                       "model.cfg.max_position_embeddings=32", "model.cfg.pretrained_model_weight=null"])
    ids, am = torch.randint(1, 128, (4, 2, 16)), torch.ones(4, 2, 16, dtype=torch.long)
    assert Simcse_unsup(c.model.cfg).eval()(ids, am, labels=torch.tensor([1, 2, 3, 4]))["sim"].shape == (4,)
    c.model.cfg.pooler_type = "first-last-avg"
    out = Simcse_unsup(c.model.cfg).train()(ids, am)
    out["loss"].backward()
    assert torch.isfinite(out["loss"])


def test_couplets_and_qqp(tmp_path):
    from libai_b200.data.structures import Instance
    from projects.QQP.dataset.qqp_dataset import QQPDataset
    from projects.QQP.tokenizer.tokenizer import _BertCNWWMTokenizer

    d = str(tmp_path)
    for sp in ("train", "test"):
        os.makedirs(f"{d}/{sp}")
        open(f"{d}/{sp}/in.txt", "w").write("天 增 岁 月\n春 满 乾 坤\n")
        open(f"{d}/{sp}/out.txt", "w").write("春 满 乾 坤\n天 增 岁 月\n")
    open(f"{d}/vocabs", "w").write("\n".join(["<pad>", "<unk>", "<bos>", "<eos>", "天", "增", "岁", "月", "春", "满", "乾", "坤"]) + "\n")
    c = LazyConfig.load("projects/Couplets/configs/config.py")
    c = LazyConfig.apply_overrides(c, [f"dataloader.train.dataset.0.path={d}", "dataloader.train.dataset.0.maxlen=8",
                                       "model.cfg.vocab_size=16", "model.cfg.max_position_embeddings=8", "model.cfg.hidden_size=32",
                                       "model.cfg.intermediate_size=32", "model.cfg.hidden_layers=2", "model.cfg.num_attention_heads=4"])
    ds = instantiate(c.dataloader.train.dataset[0])
    b = Instance.stack([ds[0], ds[1]])
    out = instantiate(c.model).train()(**{k: v.tensor for k, v in b.get_fields().items()})
    out["total_loss"].backward()
    assert torch.isfinite(out["total_loss"])

    assert LazyConfig.load("projects/QQP/configs/config_qqp.py").model.cfg.num_classes == 2
    open(f"{d}/train.tsv", "w").write("id\tqid1\tqid2\tquestion1\tquestion2\tis_duplicate\n0\t1\t2\t天 增\t岁 月\t1\n")
    open(f"{d}/vocab.txt", "w").write("\n".join(["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]", "[BOS]", "[EOS]", "天", "增", "岁", "月"]) + "\n")
    q = QQPDataset("t", [f"{d}/train.tsv"], _BertCNWWMTokenizer(f"{d}/vocab.txt"), 12)
    assert q[0].model_input.tensor.tolist()[:7] == [2, 7, 8, 3, 9, 10, 3] and int(q[0].labels.tensor) == 1


def test_chatglm_lora_loader_roundtrip(tmp_path):
    """ChatGLMLoraLoaderLiBai: dense checkpoint → LoRA-wrapped model → adapter weights restored from a LoRA checkpoint."""
    from libai_b200.utils.checkpoint import Checkpointer
    from projects.ChatGLM.chatglm import ChatGLMForConditionalGeneration
    from projects.ChatGLM.lora.lora_model import LoraModel
    from projects.ChatGLM.utils.chatglm_loader import ChatGLMLoraLoaderLiBai

    c = LazyConfig.load("projects/ChatGLM/configs/chatglm_config.py")
    c = LazyConfig.apply_overrides(c, ["cfg.num_layers=2", "cfg.hidden_size=64", "cfg.ffn_hidden_size=96", "cfg.num_attention_heads=4",
                                       "cfg.kv_channels=16", "cfg.multi_query_group_num=2", "cfg.padded_vocab_size=128",
                                       "cfg.seq_length=32"])
    lora_cfg = DictConfig(dict(r=4, lora_alpha=8, lora_dropout=0.0, target_modules=["query_key_value"], fan_in_fan_out=False,
                               bias="none", modules_to_save=None))
    torch.manual_seed(0)
    dense = ChatGLMForConditionalGeneration(c.cfg)
    Checkpointer(dense, str(tmp_path / "dense")).save("model_final")
    tuned = ChatGLMForConditionalGeneration(c.cfg)
    tuned.load_state_dict(dense.state_dict())
    tuned.transformer = LoraModel(tuned.transformer, lora_cfg, "default")
    with torch.no_grad():
        for n, p in tuned.named_parameters():
            if "lora_B" in n:
                p.normal_(std=0.05)
    torch.save(tuned.transformer.state_dict(), tmp_path / "lora.pt")        # superset of the adapter tensors
    model, info = ChatGLMLoraLoaderLiBai(ChatGLMForConditionalGeneration, c.cfg, str(tmp_path / "dense"), lora_cfg=lora_cfg,
                                         lora_pretrained_model_path=str(tmp_path / "lora.pt"), output_loading_info=True).load()
    assert not info["missing_keys"] and not info["mismatched_keys"]
    ids = torch.randint(1, 128, (2, 10))
    model.eval(), tuned.eval(), dense.eval()
    with torch.no_grad():
        a, b, d = (m(ids)["logits"] if "logits" in m(ids) else list(m(ids).values())[0] for m in (model, tuned, dense))
    assert torch.allclose(a, b, atol=1e-5) and not torch.allclose(a, d, atol=1e-4)
    fresh = ChatGLMLoraLoaderLiBai(ChatGLMForConditionalGeneration, c.cfg, str(tmp_path / "dense"), lora_cfg=lora_cfg).load()
    with torch.no_grad():
        f = fresh.eval()(ids)
    assert torch.allclose(f["logits"] if "logits" in f else list(f.values())[0], d, atol=1e-5)   # lora_B starts at zero


def test_couplet_pipeline(tmp_path):
    """projects/Couplets/distribute_infer.py: BasePipeline subclass == the single-process generator on the same weights."""
    from libai_b200.utils import distributed as dist
    from libai_b200.utils.checkpoint import Checkpointer
    from projects.Couplets.distribute_infer import CoupletPipeline
    from projects.Couplets.infer import GeneratorForEager

    vocab = tmp_path / "vocab.txt"
    vocab.write_text("\n".join(["<pad>", "<unk>", "<bos>", "<eos>", "天", "增", "岁", "月", "春", "满", "乾", "坤"]) + "\n")
    over = ["model.cfg.vocab_size=16", "model.cfg.max_position_embeddings=16", "model.cfg.hidden_size=32",
            "model.cfg.intermediate_size=32", "model.cfg.hidden_layers=2", "model.cfg.num_attention_heads=4"]
    cfg = LazyConfig.apply_overrides(LazyConfig.load("projects/Couplets/configs/config.py"), over)
    dist.reset_dist_util()
    try:
        pipe = CoupletPipeline(cfg, data_parallel=1, tensor_parallel=1, pipeline_parallel=1, mode="random", device="cpu",
                               vocab_file=str(vocab))
        out = pipe("天增岁月")
        assert set(out) == {"generated_text"} and isinstance(out["generated_text"], str)
        assert pipe.generate("天增岁月") == out["generated_text"]                   # greedy: deterministic
        Checkpointer(pipe.model, str(tmp_path / "ckpt")).save("model_final")
        dist.reset_dist_util()
        again = CoupletPipeline(LazyConfig.apply_overrides(LazyConfig.load("projects/Couplets/configs/config.py"), over),
                                1, 1, 1, model_path=str(tmp_path / "ckpt" / "model_final"), device="cpu", vocab_file=str(vocab))
        assert again("天增岁月") == out
    finally:
        dist.reset_dist_util()
