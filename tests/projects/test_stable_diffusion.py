"""Stable Diffusion project: module shapes / parameter counts, schedules, LoRA, checkpoint folder round trip,
datasets and the sampler, all on tiny random models."""
import numpy as np
import torch

from libai_b200.config import DictConfig, LazyConfig
from libai_b200.utils import distributed as dist


def _setup():
    dist.setup_dist_util(DictConfig(dict(data_parallel_size=1, tensor_parallel_size=1, pipeline_parallel_size=1,
                                         device_type="cpu")))


class FakeTok:
    model_max_length = 16

    def __call__(self, text, padding=None, truncation=None, max_length=16, return_tensors=None):
        texts = [text] if isinstance(text, str) else text
        ids = [[(ord(c) % 900) + 1 for c in t][:max_length] for t in texts]
        ids = [i + [0] * (max_length - len(i)) for i in ids]
        out = type("Enc", (), {})()
        out.input_ids = torch.tensor(ids) if return_tensors == "pt" else (ids[0] if isinstance(text, str) else ids)
        return out


def test_reference_sizes_match_released_checkpoints():
    from projects.Stable_Diffusion.modules.unet import UNet2DConditionModel
    from projects.Stable_Diffusion.modules.vae import AutoencoderKL

    _setup()
    with torch.device("meta"):
        unet, vae = UNet2DConditionModel(), AutoencoderKL()
    assert sum(p.numel() for p in unet.parameters()) == 859_520_964          # SD v1.x UNet
    assert sum(p.numel() for p in vae.parameters()) == 83_653_863            # SD v1.x VAE
    keys = set(unet.state_dict())
    for k in ("conv_in.weight", "time_embedding.linear_1.weight", "down_blocks.0.resnets.0.time_emb_proj.weight",
              "down_blocks.0.attentions.1.transformer_blocks.0.attn2.to_k.weight",
              "down_blocks.2.downsamplers.0.conv.weight", "mid_block.attentions.0.proj_in.weight",
              "up_blocks.1.attentions.2.transformer_blocks.0.ff.net.0.proj.weight",
              "up_blocks.0.upsamplers.0.conv.bias", "up_blocks.3.resnets.2.conv_shortcut.weight", "conv_out.bias"):
        assert k in keys, k
    assert unet.state_dict()["up_blocks.3.resnets.0.conv1.weight"].shape == (320, 960, 3, 3)
    vk = set(vae.state_dict())
    for k in ("encoder.down_blocks.3.resnets.1.conv2.weight", "encoder.mid_block.attentions.0.to_q.weight",
              "decoder.up_blocks.0.upsamplers.0.conv.weight", "quant_conv.weight", "post_quant_conv.bias"):
        assert k in vk, k


def test_scheduler_forward_and_ddim_inverse():
    from projects.Stable_Diffusion.modules.scheduler import DDPMScheduler

    s = DDPMScheduler()
    assert abs(float(s.alphas_cumprod[-1]) - 0.00466) < 2e-4 and abs(float(s.betas[0]) - 0.00085) < 1e-6
    x0, eps = torch.randn(3, 4, 8, 8), torch.randn(3, 4, 8, 8)
    t = torch.tensor([10, 500, 999])
    xt = s.add_noise(x0, eps, t)
    a = s.alphas_cumprod[t].view(-1, 1, 1, 1)
    assert torch.allclose(xt, a.sqrt() * x0 + (1 - a).sqrt() * eps, atol=1e-6)
    v = s.get_velocity(x0, eps, t)
    assert torch.allclose(a.sqrt() * xt - (1 - a).sqrt() * v, x0, atol=1e-4)          # v-parameterisation identity
    # with the true noise as "model output", deterministic DDIM walks back to x0 exactly
    s.set_timesteps(20)
    x = s.add_noise(x0[:1], eps[:1], s.timesteps[:1])
    for ts in s.timesteps:
        x = s.step_ddim(eps[:1], ts, x)
    assert (x - x0[:1]).abs().max() < 0.1


def test_training_loss_lora_and_prior_preservation():
    from projects.Stable_Diffusion.modeling import StableDiffusion

    _setup()
    torch.manual_seed(0)
    px, ids = torch.randn(4, 3, 32, 32), torch.randint(0, 1000, (4, 16))
    m = StableDiffusion(tiny=True).train()
    out = m(px, ids)
    out["loss"].backward()
    assert 0.3 < float(out["loss"]) < 3.0
    assert all(p.grad is None for p in m.vae.parameters()) and not m.vae.training
    assert all(p.grad is not None for p in m.unet.parameters())

    lora = StableDiffusion(tiny=True, train_with_lora=True, with_prior_preservation=True, prior_loss_weight=0.5).train()
    trainable = [n for n, p in lora.named_parameters() if p.requires_grad]
    assert trainable and all("lora" in n for n in trainable)
    StableDiffusion.set_activation_checkpoint(lora)
    loss = lora(px, ids)["loss"]
    loss.backward()
    ups = [p for n, p in lora.named_parameters() if n.endswith("to_q_lora.down.weight")]
    assert all(p.grad is not None for p in ups)          # (up is zero-initialised → only `up` gets a non-zero grad)
    assert any(p.grad.abs().sum() > 0 for n, p in lora.named_parameters() if n.endswith("to_q_lora.up.weight"))


def test_pipeline_save_load_and_lora_roundtrip(tmp_path):
    from projects.Stable_Diffusion.modeling import StableDiffusion
    from projects.Stable_Diffusion.modules.loader import load_submodel
    from projects.Stable_Diffusion.modules.lora import load_attn_procs, save_attn_procs
    from projects.Stable_Diffusion.modules.unet import UNet2DConditionModel
    from projects.Stable_Diffusion.pipeline import StableDiffusionPipeline

    _setup()
    torch.manual_seed(0)
    m = StableDiffusion(tiny=True, train_with_lora=True).eval()
    for p in m.lora_layers.parameters():
        torch.nn.init.normal_(p, std=0.05)
    pipe = StableDiffusionPipeline(None, m.text_encoder, m.vae, m.unet, m.noise_scheduler)
    ids = torch.randint(1, 1000, (1, 16))
    g = torch.Generator().manual_seed(1)
    img = pipe(input_ids=ids, height=16, width=16, num_inference_steps=3, guidance_scale=2.0, generator=g, output_type="pt")
    assert img.shape == (1, 3, 16, 16) and 0 <= float(img.min()) and float(img.max()) <= 1

    pipe.save_pretrained(str(tmp_path / "sd"))
    save_attn_procs(m.lora_layers, str(tmp_path / "lora"))
    unet2 = load_submodel(UNet2DConditionModel, str(tmp_path / "sd"), "unet")
    load_attn_procs(unet2, str(tmp_path / "lora"))
    x, t, ctx = torch.randn(1, 4, 4, 4), torch.tensor([7]), torch.randn(1, 16, 32)
    with torch.no_grad():
        assert torch.allclose(m.unet(x, t, ctx).sample, unet2.eval()(x, t, ctx).sample, atol=1e-5)
    pil = pipe(input_ids=ids, height=16, width=16, num_inference_steps=2, guidance_scale=1.0)
    assert pil[0].size == (16, 16)


def test_datasets(tmp_path):
    from PIL import Image

    from projects.Stable_Diffusion.dataset import (
        DreamBoothDataset, PromptDataset, TXTDataset, prior_preservation_collate,
    )

    rng = np.random.default_rng(0)
    for sub, n in (("inst", 2), ("cls", 3), ("txt", 2)):
        (tmp_path / sub).mkdir()
        for i in range(n):
            Image.fromarray(rng.integers(0, 255, (40, 48, 3), dtype=np.uint8)).save(tmp_path / sub / f"{i}.jpg")
            if sub == "txt":
                (tmp_path / sub / f"{i}.txt").write_text(f"caption {i}")
    ds = DreamBoothDataset(str(tmp_path / "inst"), "a photo of sks dog", FakeTok(), class_data_root=str(tmp_path / "cls"),
                           class_prompt="a photo of dog", size=32)
    assert len(ds) == 3
    batch = prior_preservation_collate([ds[0], ds[1]])
    assert batch.get("pixel_values").tensor.shape == (4, 3, 32, 32) and batch.get("input_ids").tensor.shape == (4, 16)
    assert torch.equal(batch.get("input_ids").tensor[0], batch.get("input_ids").tensor[1])      # instance half first
    assert not torch.equal(batch.get("input_ids").tensor[0], batch.get("input_ids").tensor[2])
    txt = TXTDataset(str(tmp_path / "txt"), FakeTok(), size=32)
    assert len(txt) == 2 and txt[0].get("pixel_values").tensor.shape == (3, 32, 32)
    assert float(txt[0].get("pixel_values").tensor.min()) >= -1.0
    assert PromptDataset("p", 5)[3] == {"prompt": "p", "index": 3}


def test_configs_load():
    for name in ("config", "dreambooth_config", "prior_preservation_config", "lora_config"):
        cfg = LazyConfig.load(f"projects/Stable_Diffusion/configs/{name}.py")
        assert cfg.model.model_path and cfg.train.train_iter > 0
    assert LazyConfig.load("projects/Stable_Diffusion/configs/lora_config.py").model.train_with_lora is True
    assert LazyConfig.load("projects/Stable_Diffusion/configs/prior_preservation_config.py").model.with_prior_preservation
