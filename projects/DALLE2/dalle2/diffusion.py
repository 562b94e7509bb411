"""Gaussian diffusion bookkeeping shared by the DALL-E 2 prior and decoder.

Spec: reference projects/DALLE2/dalle2/models.py:248-395 (beta schedules, ``NoiseScheduler`` with ``q_sample``,
``q_posterior``, ``predict_start_from_noise``, loss and p2 re-weighting) — the standard DDPM algebra
(Ho et al. 2020; Nichol & Dhariwal 2021)."""
import math

import torch
import torch.nn.functional as F
from torch import nn


def cosine_beta_schedule(timesteps, s=0.008):
    x = torch.linspace(0, timesteps, timesteps + 1, dtype=torch.float64)
    bar = torch.cos(((x / timesteps) + s) / (1 + s) * math.pi * 0.5) ** 2
    bar = bar / bar[0]
    return (1 - bar[1:] / bar[:-1]).clamp(0, 0.999)


def linear_beta_schedule(timesteps):
    scale = 1000 / timesteps
    return torch.linspace(scale * 1e-4, scale * 0.02, timesteps, dtype=torch.float64)


def quadratic_beta_schedule(timesteps):
    scale = 1000 / timesteps
    return torch.linspace((scale * 1e-4) ** 0.5, (scale * 0.02) ** 0.5, timesteps, dtype=torch.float64) ** 2


def sigmoid_beta_schedule(timesteps):
    scale = 1000 / timesteps
    return torch.sigmoid(torch.linspace(-6, 6, timesteps, dtype=torch.float64)) * (scale * 0.02 - scale * 1e-4) + scale * 1e-4


_SCHEDULES = {"cosine": cosine_beta_schedule, "linear": linear_beta_schedule, "quadratic": quadratic_beta_schedule,
              "jsd": lambda t: 1.0 / torch.linspace(t, 1, t, dtype=torch.float64), "sigmoid": sigmoid_beta_schedule}


def extract(a, t, x_shape):
    return a.gather(-1, t).reshape(t.shape[0], *((1,) * (len(x_shape) - 1)))


def meanflat(x):
    return x.mean(dim=tuple(range(1, x.dim())))


def normal_kl(mean1, logvar1, mean2, logvar2):
    return 0.5 * (-1.0 + logvar2 - logvar1 + torch.exp(logvar1 - logvar2) + (mean1 - mean2) ** 2 * torch.exp(-logvar2))


def approx_standard_normal_cdf(x):
    return 0.5 * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * x ** 3)))


def discretized_gaussian_log_likelihood(x, *, means, log_scales, thres=0.999):
    """Log-likelihood of 8-bit image data ``x`` ∈ [-1, 1] under a Gaussian discretised to 1/255 bins."""
    centered = x - means
    inv_std = torch.exp(-log_scales)
    cdf_plus = approx_standard_normal_cdf(inv_std * (centered + 1.0 / 255.0))
    cdf_min = approx_standard_normal_cdf(inv_std * (centered - 1.0 / 255.0))
    log_cdf_plus = torch.log(cdf_plus.clamp(min=1e-12))
    log_one_minus_cdf_min = torch.log((1.0 - cdf_min).clamp(min=1e-12))
    log_delta = torch.log((cdf_plus - cdf_min).clamp(min=1e-12))
    return torch.where(x < -thres, log_cdf_plus, torch.where(x > thres, log_one_minus_cdf_min, log_delta))


class NoiseScheduler(nn.Module):
    def __init__(self, *, beta_schedule, timesteps, loss_type="l2", p2_loss_weight_gamma=0.0, p2_loss_weight_k=1):
        super().__init__()
        assert beta_schedule in _SCHEDULES, beta_schedule
        # (the 1000/T rescaling of the non-cosine schedules exceeds 1 for very short chains: clamp like the cosine one)
        betas = _SCHEDULES[beta_schedule](timesteps).clamp(max=0.999)
        alphas = 1.0 - betas
        bar = torch.cumprod(alphas, dim=0)
        bar_prev = F.pad(bar[:-1], (1, 0), value=1.0)
        self.num_timesteps = int(timesteps)
        self.loss_type = loss_type
        self.loss_fn = {"l1": F.l1_loss, "l2": F.mse_loss, "huber": F.smooth_l1_loss}[loss_type]

        def reg(name, v):
            self.register_buffer(name, v.float(), persistent=False)

        reg("betas", betas)
        reg("alphas_cumprod", bar)
        reg("alphas_cumprod_prev", bar_prev)
        reg("sqrt_alphas_cumprod", bar.sqrt())
        reg("sqrt_one_minus_alphas_cumprod", (1 - bar).sqrt())
        reg("log_one_minus_alphas_cumprod", (1 - bar).log())
        reg("sqrt_recip_alphas_cumprod", (1 / bar).sqrt())
        reg("sqrt_recipm1_alphas_cumprod", (1 / bar - 1).sqrt())
        post_var = betas * (1 - bar_prev) / (1 - bar)
        reg("posterior_variance", post_var)
        reg("posterior_log_variance_clipped", post_var.clamp(min=1e-20).log())
        reg("posterior_mean_coef1", betas * bar_prev.sqrt() / (1 - bar))
        reg("posterior_mean_coef2", (1 - bar_prev) * alphas.sqrt() / (1 - bar))
        self.has_p2_loss_reweighting = p2_loss_weight_gamma > 0
        reg("p2_loss_weight", (p2_loss_weight_k + bar / (1 - bar)) ** -p2_loss_weight_gamma)

    def sample_random_times(self, batch, device=None):
        return torch.randint(0, self.num_timesteps, (batch,), device=device or self.betas.device, dtype=torch.long)

    def q_posterior(self, x_start, x_t, t):
        mean = extract(self.posterior_mean_coef1, t, x_t.shape) * x_start + extract(self.posterior_mean_coef2, t, x_t.shape) * x_t
        return mean, extract(self.posterior_variance, t, x_t.shape), extract(self.posterior_log_variance_clipped, t, x_t.shape)

    def q_sample(self, x_start, t, noise=None):
        noise = torch.randn_like(x_start) if noise is None else noise
        return extract(self.sqrt_alphas_cumprod, t, x_start.shape) * x_start + \
            extract(self.sqrt_one_minus_alphas_cumprod, t, x_start.shape) * noise

    def predict_start_from_noise(self, x_t, t, noise):
        return extract(self.sqrt_recip_alphas_cumprod, t, x_t.shape) * x_t - \
            extract(self.sqrt_recipm1_alphas_cumprod, t, x_t.shape) * noise

    def predict_noise_from_start(self, x_t, t, x0):
        return (extract(self.sqrt_recip_alphas_cumprod, t, x_t.shape) * x_t - x0) / \
            extract(self.sqrt_recipm1_alphas_cumprod, t, x_t.shape)

    def p2_reweigh_loss(self, loss, times):
        return loss * extract(self.p2_loss_weight, times, loss.shape) if self.has_p2_loss_reweighting else loss
