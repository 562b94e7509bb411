"""4× / 16× upsampling of generated images with SwinIR-L real-SR (reference projects/DALLE2/swinir/upsample.py)."""
import os

import torch

from .models import SwinIR

REAL_SR_LARGE = dict(upscale=4, in_chans=3, img_size=64, window_size=8, img_range=1.0, depths=[6] * 9, embed_dim=240,
                     num_heads=[8] * 9, mlp_ratio=2, upsampler="nearest+conv", resi_connection="3conv")


def load_model(model_path=None, **overrides):
    """SwinIR-L x4 GAN model; ``model_path``: the public ``003_realSR_BSRGAN_DFOWMFC_s64w8_SwinIR-L_x4_GAN.pth``
    (there is no network access here: the file must already exist; without it the weights stay random)."""
    model = SwinIR(**{**REAL_SR_LARGE, **overrides})
    if model_path and os.path.exists(model_path):
        state = torch.load(model_path, map_location="cpu", weights_only=True)
        model.load_state_dict(state.get("params_ema", state), strict=True)
    elif model_path:
        raise FileNotFoundError(f"{model_path} not found — download it from the SwinIR v0.0 release page")
    return model.eval()


@torch.no_grad()
def upsample4x(img_lq, model):
    device = next(model.parameters()).device
    return model(img_lq.to(device=device, dtype=next(model.parameters()).dtype)).float().clamp_(0, 1)


def upsample16x(imgs, model):
    return upsample4x(upsample4x(imgs, model), model)
