"""DALL-E 2 project on tiny random models: diffusion algebra, prior/decoder training losses and sampling, the
full text→image path, SwinIR, and tensor-parallel equivalence of the prior transformer."""
import torch

from libai_b200.config import DictConfig, LazyConfig
from libai_b200.utils import distributed as dist
from tests.dist_utils import run_distributed


def _setup(tp=1):
    dist.setup_dist_util(DictConfig(dict(data_parallel_size=1, tensor_parallel_size=tp, pipeline_parallel_size=1,
                                         device_type="cpu")))


def _tiny_clip():
    from projects.CLIP.clip.model import CLIP
    from projects.DALLE2.dalle2 import OpenAIClipAdapter

    torch.manual_seed(0)
    clip = CLIP(embed_dim=32, image_resolution=32, vision_layers=2, vision_width=64, vision_patch_size=8,
                context_length=12, vocab_size=49408, transformer_width=32, transformer_heads=2, transformer_layers=2)
    return OpenAIClipAdapter(clip=clip.float())


def _tiny_dalle2():
    from projects.DALLE2.dalle2 import DALLE2, Decoder, DiffusionPrior, DiffusionPriorNetwork, Unet

    clip = _tiny_clip()
    net = DiffusionPriorNetwork(dim=32, depth=2, num_timesteps=10, max_text_len=12, dim_head=8, heads=4, ff_mult=2,
                                normformer=True)
    prior = DiffusionPrior(net, clip=clip, image_embed_dim=32, timesteps=10, cond_drop_prob=0.1)
    unet = Unet(dim=16, image_embed_dim=32, text_embed_dim=32, cond_dim=16, dim_mults=(1, 2), num_resnet_blocks=1,
                attn_heads=2, attn_dim_head=8, cond_on_text_encodings=True, self_attn=[False, True], max_text_len=12,
                init_cross_embed_kernel_sizes=(3, 5))
    unet2 = Unet(dim=8, image_embed_dim=32, dim_mults=(1, 2), num_resnet_blocks=1, attn_heads=2, attn_dim_head=8,
                 init_cross_embed_kernel_sizes=(3, 5))
    decoder = Decoder((unet, unet2), clip=clip, image_sizes=[16, 32], timesteps=6, learned_variance=[True, False],
                      beta_schedule=["cosine", "linear"])
    return DALLE2(prior=prior, decoder=decoder, prior_num_samples=2), clip


def test_noise_scheduler_identities():
    from projects.DALLE2.dalle2.diffusion import NoiseScheduler

    for sched in ("cosine", "linear", "quadratic", "sigmoid"):
        ns = NoiseScheduler(beta_schedule=sched, timesteps=50)
        x0, eps = torch.randn(4, 3, 8, 8), torch.randn(4, 3, 8, 8)
        t = torch.tensor([0, 10, 30, 49])
        xt = ns.q_sample(x0, t, eps)
        assert torch.allclose(ns.predict_start_from_noise(xt, t, eps), x0, atol=2e-3)
        assert torch.allclose(ns.predict_noise_from_start(xt, t, x0), eps, atol=2e-2)
        mean, var, logvar = ns.q_posterior(x0, xt, t)
        assert mean.shape == xt.shape and (var >= 0).all() and torch.isfinite(logvar).all()
        assert ns.alphas_cumprod[0] > ns.alphas_cumprod[-1] > 0


def test_prior_and_decoder_losses_and_sampling():
    _setup()
    model, clip = _tiny_dalle2()
    text = torch.randint(1, 49000, (2, 12))
    text[:, -1] = 49407
    image = torch.rand(2, 3, 32, 32)
    model.train()
    lp = model.prior(text=text, image=image)
    lp.backward()
    assert torch.isfinite(lp) and lp > 0
    assert all(p.grad is None for p in clip.parameters())
    for n in (1, 2):
        ld = model.decoder(image, text=text, unet_number=n)
        ld.backward()
        assert torch.isfinite(ld) and ld > 0
    assert model.decoder.unets[0].channels_out == 6 and model.decoder.unets[1].lowres_cond

    torch.manual_seed(0)
    out = model(text, cond_scale=2.0, prior_cond_scale=1.5)
    assert out.shape == (2, 3, 32, 32) and 0 <= float(out.min()) and float(out.max()) <= 1
    pil = model(text[:1], return_pil_images=True)
    assert pil.size == (32, 32)
    emb = model.prior.sample(text, num_samples_per_batch=3)
    assert emb.shape == (2, 32)


def test_swinir_shapes_and_checkpoint_keys():
    _setup()
    from projects.DALLE2.swinir import SwinIR, upsample4x

    torch.manual_seed(0)
    m = SwinIR(upscale=4, window_size=4, depths=[2, 2], embed_dim=24, num_heads=[2, 2], mlp_ratio=2,
               upsampler="nearest+conv", resi_connection="3conv").eval()
    y = upsample4x(torch.rand(1, 3, 10, 14), m)
    assert y.shape == (1, 3, 40, 56)
    keys = set(m.state_dict())
    for k in ("conv_first.weight", "patch_embed.norm.weight", "layers.0.residual_group.blocks.1.attn.qkv.weight",
              "layers.1.residual_group.blocks.0.attn.relative_position_bias_table",
              "layers.0.residual_group.blocks.0.mlp.fc1.weight", "layers.0.conv.4.weight", "norm.bias",
              "conv_after_body.0.weight", "conv_before_upsample.0.weight", "conv_up2.bias", "conv_hr.weight",
              "conv_last.weight"):
        assert k in keys, k
    for up in ("pixelshuffle", "pixelshuffledirect", ""):
        net = SwinIR(upscale=2 if up else 1, window_size=4, depths=[2], embed_dim=12, num_heads=[2], upsampler=up).eval()
        with torch.no_grad():
            assert net(torch.rand(1, 3, 8, 8)).shape[-1] == (16 if up else 8)


def _tp_prior(rank, world):
    from projects.DALLE2.dalle2.dalle2_loader import load_state
    from projects.DALLE2.dalle2.prior import DiffusionPriorNetwork

    kw = dict(dim=32, depth=2, num_timesteps=10, max_text_len=6, dim_head=8, heads=4, ff_mult=2, normformer=True)
    x, t = torch.randn(3, 32, generator=torch.Generator().manual_seed(1)), torch.tensor([1, 5, 9])
    te = torch.randn(3, 32, generator=torch.Generator().manual_seed(2))
    enc = torch.randn(3, 6, 32, generator=torch.Generator().manual_seed(3))
    _setup(1)
    torch.manual_seed(0)
    full = DiffusionPriorNetwork(**kw).eval()
    with torch.no_grad():
        for p in full.parameters():
            p.copy_(torch.randn(p.shape, generator=torch.Generator().manual_seed(p.numel())) * 0.1)
        want = full(x, t, text_embed=te, text_encodings=enc)
    state = {k: v.clone() for k, v in full.state_dict().items()}
    _setup(world)
    shard = DiffusionPriorNetwork(**kw).eval()
    own = dict(shard.named_parameters())
    with torch.no_grad():
        for k, v in state.items():
            if k not in own:
                continue
            tp_dim = getattr(own[k], "tp_dim", None)
            if k.endswith("post_norm.weight") or k.endswith("post_norm.bias"):
                tp_dim = 0
            own[k].copy_(v if tp_dim is None else v.chunk(world, dim=tp_dim)[rank])
        got = shard(x, t, text_embed=te, text_encodings=enc)
    return float((got - want).abs().max())


def test_prior_tensor_parallel_matches_single():
    for err in run_distributed(_tp_prior, 2):
        assert err < 1e-4, err


def test_config_and_loader_translation():
    cfg = LazyConfig.load("projects/DALLE2/configs/dalle2_config.py")
    assert cfg.prior.net.depth == 24 and cfg.decoder.learned_variance is True and cfg.unet1.dim == 320
    from projects.DALLE2.dalle2.dalle2_loader import _translate

    assert _translate("net.causal_transformer.layers.3.0.to_out.0.weight") == "net.causal_transformer.layers.3.0.to_out.weight"
    assert _translate("net.causal_transformer.layers.3.1.5.weight") == "net.causal_transformer.layers.3.1.w_out.weight"


def test_inference_pipeline(tmp_path):
    from libai_b200.config import LazyCall
    from projects.DALLE2.dalle2_inference import Dalle2Pipeline

    cfg = LazyConfig.load("projects/DALLE2/configs/dalle2_config.py")
    cfg.clip = LazyCall(_tiny_clip)()
    cfg.prior.clip = cfg.clip
    cfg.prior.image_embed_dim = 32
    cfg.prior.timesteps = 4
    cfg.prior.net.update(dict(dim=32, depth=1, num_timesteps=4, max_text_len=12, dim_head=8, heads=2, ff_mult=2))
    cfg.unet1.update(dict(dim=16, image_embed_dim=32, text_embed_dim=32, cond_dim=16, dim_mults=(1, 2), num_resnet_blocks=1,
                          attn_heads=2, attn_dim_head=8, self_attn=[False, True], max_text_len=12,
                          init_cross_embed_kernel_sizes=(3, 5)))
    cfg.decoder.unet = (cfg.unet1,)
    cfg.decoder.image_sizes = [16]
    cfg.decoder.timesteps = 3
    cfg.model.prior, cfg.model.decoder = cfg.prior, cfg.decoder
    pipe = Dalle2Pipeline(cfg, data_parallel=1, tensor_parallel=1, pipeline_parallel=1, mode="random", device="cpu",
                          save_images=True, output_dir=str(tmp_path))
    pipe.tokenizer = type("T", (), {"tokenize": staticmethod(lambda texts: torch.cat(
        [torch.randint(1, 49000, (len(texts), 11)), torch.full((len(texts), 1), 49407)], dim=1))})()
    out = pipe(["a shiba inu wearing a beret", "a teddy bear on a skateboard"])
    assert out["image_embed"].shape == (2, 3, 16, 16)
    assert len(list(tmp_path.glob("*.png"))) == 2
