"""Noise schedules for training (DDPM forward process) and sampling (DDPM ancestral and DDIM steps).

Covers what the reference uses from ``diffusers.DDPMScheduler`` (projects/Stable_Diffusion/modeling.py:42,107-129:
``add_noise``, ``get_velocity``, ``config.prediction_type``, ``config.num_train_timesteps``) plus the sampler side
needed by ``pipeline.py``.  The schedule tensors are kept on the device of the latents so ``add_noise`` is a pure
device op (no host sync inside the training step)."""
import json
import os

import torch


class SchedulerConfig(dict):
    __getattr__ = dict.__getitem__


class DDPMScheduler:
    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                 prediction_type="epsilon", clip_sample=False, set_alpha_to_one=False, steps_offset=1, **unused):
        self.config = SchedulerConfig(num_train_timesteps=num_train_timesteps, beta_start=beta_start,
                                      beta_end=beta_end, beta_schedule=beta_schedule,
                                      prediction_type=prediction_type, clip_sample=clip_sample,
                                      set_alpha_to_one=set_alpha_to_one, steps_offset=steps_offset)
        if beta_schedule == "linear":
            betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float64)
        elif beta_schedule == "scaled_linear":
            betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float64) ** 2
        elif beta_schedule == "squaredcos_cap_v2":
            t = torch.arange(num_train_timesteps + 1, dtype=torch.float64) / num_train_timesteps
            bar = torch.cos((t + 0.008) / 1.008 * torch.pi / 2) ** 2
            betas = (1 - bar[1:] / bar[:-1]).clamp(max=0.999)
        else:
            raise NotImplementedError(beta_schedule)
        self.betas = betas.float()
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0).float()
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.timesteps = torch.arange(num_train_timesteps - 1, -1, -1)
        self.num_inference_steps = None

    @classmethod
    def from_pretrained(cls, model_path, subfolder="scheduler"):
        with open(os.path.join(model_path, subfolder, "scheduler_config.json")) as f:
            cfg = json.load(f)
        return cls(**{k: v for k, v in cfg.items() if not k.startswith("_")})

    def to(self, device):
        for name in ("betas", "alphas", "alphas_cumprod", "final_alpha_cumprod"):
            setattr(self, name, getattr(self, name).to(device))
        return self

    def _coeffs(self, timesteps, like):
        if self.alphas_cumprod.device != like.device:
            self.to(like.device)
        a = self.alphas_cumprod[timesteps].to(like.dtype)
        shape = (-1,) + (1,) * (like.dim() - 1)
        return a.sqrt().view(shape), (1 - a).sqrt().view(shape)

    def add_noise(self, original, noise, timesteps):
        sa, sb = self._coeffs(timesteps, original)
        return sa * original + sb * noise

    def get_velocity(self, sample, noise, timesteps):
        sa, sb = self._coeffs(timesteps, sample)
        return sa * noise - sb * sample

    # ------------------------------------------------------------------ sampling
    def set_timesteps(self, num_inference_steps, device=None):
        self.num_inference_steps = num_inference_steps
        ratio = self.config.num_train_timesteps // num_inference_steps
        ts = (torch.arange(num_inference_steps) * ratio).flip(0) + self.config.steps_offset
        self.timesteps = ts.clamp(max=self.config.num_train_timesteps - 1).to(device or "cpu")
        if device is not None:
            self.to(device)

    def _x0_eps(self, model_output, t, sample):
        a = self.alphas_cumprod[t]
        if self.config.prediction_type == "epsilon":
            eps = model_output
            x0 = (sample - (1 - a).sqrt() * eps) / a.sqrt()
        elif self.config.prediction_type == "v_prediction":
            x0 = a.sqrt() * sample - (1 - a).sqrt() * model_output
            eps = a.sqrt() * model_output + (1 - a).sqrt() * sample
        elif self.config.prediction_type == "sample":
            x0 = model_output
            eps = (sample - a.sqrt() * x0) / (1 - a).sqrt()
        else:
            raise ValueError(f"Unknown prediction type {self.config.prediction_type}")
        if self.config.clip_sample:
            x0 = x0.clamp(-1, 1)
        return x0, eps

    def step_ddim(self, model_output, t, sample, eta=0.0, generator=None):
        """One deterministic (eta=0) or stochastic DDIM step from timestep ``t`` to ``t - stride``."""
        stride = self.config.num_train_timesteps // (self.num_inference_steps or self.config.num_train_timesteps)
        t = int(t)
        prev = t - stride
        a_prev = self.alphas_cumprod[prev] if prev >= 0 else self.final_alpha_cumprod
        a = self.alphas_cumprod[t]
        x0, eps = self._x0_eps(model_output.float(), t, sample.float())
        sigma = eta * ((1 - a_prev) / (1 - a) * (1 - a / a_prev)).clamp(min=0).sqrt()
        out = a_prev.sqrt() * x0 + (1 - a_prev - sigma ** 2).clamp(min=0).sqrt() * eps
        if eta > 0:
            out = out + sigma * torch.randn(sample.shape, generator=generator, device=sample.device)
        return out.to(sample.dtype)

    def step(self, model_output, t, sample, generator=None):
        """One DDPM ancestral step (posterior mean + fixed-small variance)."""
        stride = self.config.num_train_timesteps // (self.num_inference_steps or self.config.num_train_timesteps)
        t = int(t)
        prev = t - stride
        a = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev] if prev >= 0 else torch.ones_like(a)
        beta_t = 1 - a / a_prev
        x0, _ = self._x0_eps(model_output.float(), t, sample.float())
        mean = (a_prev.sqrt() * beta_t / (1 - a)) * x0 + ((a / a_prev).sqrt() * (1 - a_prev) / (1 - a)) * sample.float()
        if prev >= 0:
            var = ((1 - a_prev) / (1 - a) * beta_t).clamp(min=1e-20)
            mean = mean + var.sqrt() * torch.randn(sample.shape, generator=generator, device=sample.device)
        return mean.to(sample.dtype)
