"""DALL-E 2 decoder: a cascade of conditional U-Nets turning a CLIP image embedding (+ caption encodings) into
pixels, each stage at a higher resolution and conditioned on the previous stage's output.

Spec: reference projects/DALLE2/dalle2/models.py:1747-2500 — per-stage noise schedules, optional learned variance
(the U-Net emits 2×channels; the second half interpolates between β_t and the posterior variance in log space and is
trained with the variational-bound term), classifier-free guidance at sampling time, dynamic thresholding of the
predicted x₀, low-resolution conditioning with blur/noise augmentation for the super-resolution stages.
"""
import torch
import torch.nn.functional as F
from torch import nn

from .diffusion import NoiseScheduler, discretized_gaussian_log_likelihood, extract, meanflat, normal_kl
from .unet import Unet

NAT = 1.0 / 0.6931471805599453


def resize_image_to(image, size):
    if image.shape[-1] == size:
        return image
    return F.interpolate(image, size=(size, size), mode="nearest" if size > image.shape[-1] else "area")


def normalize_neg_one_to_one(img):
    return img * 2 - 1


def unnormalize_zero_to_one(img):
    return (img + 1) * 0.5


class LowresConditioner(nn.Module):
    def __init__(self, downsample_first=True, blur_sigma=0.6, blur_kernel_size=3):
        super().__init__()
        self.downsample_first, self.blur_sigma, self.blur_kernel_size = downsample_first, blur_sigma, blur_kernel_size

    def _blur(self, x):
        k = self.blur_kernel_size
        ax = torch.arange(k, device=x.device, dtype=torch.float32) - (k - 1) / 2
        g = torch.exp(-0.5 * (ax / self.blur_sigma) ** 2)
        g = (g / g.sum()).to(x.dtype)
        c = x.shape[1]
        x = F.conv2d(F.pad(x, (k // 2,) * 4, mode="reflect"), g.view(1, 1, 1, k).expand(c, 1, 1, k), groups=c)
        return F.conv2d(x, g.view(1, 1, k, 1).expand(c, 1, k, 1), groups=c)

    def forward(self, cond_fmap, *, target_image_size, downsample_image_size=None):
        if self.training and self.downsample_first and downsample_image_size is not None:
            cond_fmap = resize_image_to(cond_fmap, downsample_image_size)
        if self.training:
            cond_fmap = self._blur(cond_fmap)
        return resize_image_to(cond_fmap, target_image_size)


class Decoder(nn.Module):
    def __init__(self, unet, *, clip=None, image_size=None, channels=3, timesteps=1000, image_cond_drop_prob=0.1,
                 text_cond_drop_prob=0.5, loss_type="l2", beta_schedule=None, predict_x_start=False, image_sizes=None,
                 learned_variance=True, vb_loss_weight=0.001, clip_denoised=True, dynamic_thres_percentile=0.9,
                 use_dynamic_thres=False, lowres_downsample_first=True, blur_sigma=0.6, blur_kernel_size=3, **unused):
        super().__init__()
        self.clip = clip
        if clip is not None:
            for p in clip.parameters():
                p.requires_grad_(False)
        unets = [unet] if isinstance(unet, nn.Module) else list(unet)      # (config containers are iterable too)
        n = len(unets)
        self.channels = channels
        self.image_sizes = list(image_sizes) if image_sizes is not None else [image_size or clip.image_size]
        assert len(self.image_sizes) == n
        self.sample_channels = [channels] * n

        def per(v):
            return list(v) if hasattr(v, "__iter__") and not isinstance(v, str) else [v] * n

        self.learned_variance = per(learned_variance)
        self.predict_x_start = per(predict_x_start)
        schedules = per(beta_schedule) if beta_schedule is not None else ["cosine"] + ["linear"] * (n - 1)
        self.unets = nn.ModuleList()
        for i, (u, lv) in enumerate(zip(unets, self.learned_variance)):
            assert isinstance(u, Unet)
            want_out = channels * (2 if lv else 1)
            if u.channels_out != want_out or u.lowres_cond != (i > 0):
                u = self._rebuild(u, want_out, i > 0)
            self.unets.append(u)
        self.noise_schedulers = nn.ModuleList([NoiseScheduler(beta_schedule=s, timesteps=t, loss_type=loss_type)
                                               for s, t in zip(schedules, per(timesteps))])
        self.vb_loss_weight = vb_loss_weight
        self.to_lowres_cond = LowresConditioner(lowres_downsample_first, blur_sigma, blur_kernel_size)
        self.image_cond_drop_prob, self.text_cond_drop_prob = image_cond_drop_prob, text_cond_drop_prob
        self.can_classifier_guidance = image_cond_drop_prob > 0 or text_cond_drop_prob > 0
        self.condition_on_text_encodings = any(u.cond_on_text_encodings for u in self.unets)
        self.clip_denoised, self.use_dynamic_thres, self.dynamic_thres_percentile = clip_denoised, use_dynamic_thres, dynamic_thres_percentile

    @staticmethod
    def _rebuild(u, channels_out, lowres_cond):
        """The configured U-Net does not know yet whether it must emit a variance channel set or take a low-res
        image; swap the two affected layers in place (everything else is independent of these choices)."""
        from .unet import CrossEmbedLayer

        if u.channels_out != channels_out:
            old = u.final_conv[1]
            u.final_conv[1] = nn.Conv2d(old.in_channels, channels_out, 1).to(old.weight.device, old.weight.dtype)
            u.channels_out = channels_out
        if u.lowres_cond != lowres_cond:
            first = u.init_conv.convs
            dim_out = sum(c.out_channels for c in first)
            u.init_conv = CrossEmbedLayer(u.channels * (2 if lowres_cond else 1), [c.kernel_size[0] for c in first], dim_out)
            u.lowres_cond = lowres_cond
        return u

    # ------------------------------------------------------------------ sampling
    def _model_out(self, unet, x, t, *, image_embed, text_encodings, text_mask, lowres_cond_img, cond_scale):
        return unet.forward_with_cond_scale(x, t, image_embed=image_embed, text_encodings=text_encodings,
                                            text_mask=text_mask, lowres_cond_img=lowres_cond_img, cond_scale=cond_scale)

    def p_mean_variance(self, unet, x, t, ns, *, learned_variance, predict_x_start, **cond):
        return self._mean_variance_from_output(self._model_out(unet, x, t, **cond).float(), x, t, ns,
                                               learned_variance, predict_x_start)

    def _mean_variance_from_output(self, pred, x, t, ns, learned_variance, predict_x_start):
        if learned_variance:
            pred, var_interp = pred.chunk(2, dim=1)
        x0 = pred if predict_x_start else ns.predict_start_from_noise(x, t, pred)
        if self.clip_denoised:
            s = 1.0
            if self.use_dynamic_thres:
                s = torch.quantile(x0.flatten(1).abs(), self.dynamic_thres_percentile, dim=-1).clamp(min=1.0).view(-1, 1, 1, 1)
            x0 = x0.clamp(-s, s) / s
        mean, var, log_var = ns.q_posterior(x0, x, t)
        if learned_variance:
            min_log = extract(ns.posterior_log_variance_clipped, t, x.shape)
            max_log = extract(ns.betas.log(), t, x.shape)
            frac = unnormalize_zero_to_one(var_interp)
            log_var = frac * max_log + (1 - frac) * min_log
            var = log_var.exp()
        return mean, var, log_var

    @torch.no_grad()
    def p_sample_loop(self, unet, shape, ns, **kw):
        device = ns.betas.device
        img = torch.randn(shape, device=device)
        for i in range(ns.num_timesteps - 1, -1, -1):
            t = torch.full((shape[0],), i, device=device, dtype=torch.long)
            mean, _, log_var = self.p_mean_variance(unet, img, t, ns, **kw)
            nonzero = (t != 0).float().view(-1, 1, 1, 1)
            img = mean + nonzero * (0.5 * log_var).exp() * torch.randn_like(img)
        return unnormalize_zero_to_one(img)

    @torch.no_grad()
    def sample(self, image_embed=None, text=None, text_mask=None, text_encodings=None, batch_size=1, cond_scale=1.0,
               stop_at_unet_number=None):
        self.eval()
        if text is not None and text_encodings is None and self.clip is not None:
            _, text_encodings, text_mask = self.clip.embed_text(text)
        if image_embed is not None:
            batch_size = image_embed.shape[0]
        img = None
        for i, (unet, ns, size, lv, px0) in enumerate(zip(self.unets, self.noise_schedulers, self.image_sizes,
                                                          self.learned_variance, self.predict_x_start)):
            lowres = None
            if unet.lowres_cond:
                lowres = normalize_neg_one_to_one(self.to_lowres_cond(img, target_image_size=size))
            img = self.p_sample_loop(unet, (batch_size, self.channels, size, size), ns, learned_variance=lv,
                                     predict_x_start=px0, image_embed=image_embed,
                                     text_encodings=text_encodings if unet.cond_on_text_encodings else None,
                                     text_mask=text_mask if unet.cond_on_text_encodings else None,
                                     lowres_cond_img=lowres, cond_scale=cond_scale)
            if stop_at_unet_number is not None and stop_at_unet_number == i + 1:
                break
        return img

    # ------------------------------------------------------------------ training
    def p_losses(self, unet, x_start, times, ns, *, image_embed, text_encodings, text_mask, lowres_cond_img,
                 learned_variance, predict_x_start):
        noise = torch.randn_like(x_start)
        x_noisy = ns.q_sample(x_start, times, noise)
        out = unet(x_noisy, times, image_embed=image_embed, text_encodings=text_encodings, text_mask=text_mask,
                   lowres_cond_img=lowres_cond_img, image_cond_drop_prob=self.image_cond_drop_prob,
                   text_cond_drop_prob=self.text_cond_drop_prob).float()
        pred = out.chunk(2, dim=1)[0] if learned_variance else out
        loss = ns.loss_fn(pred, x_start if predict_x_start else noise, reduction="none")
        loss = ns.p2_reweigh_loss(meanflat(loss), times).mean()
        if not learned_variance:
            return loss
        # variational-bound term trains only the variance head (mean detached), Nichol & Dhariwal 2021 §3.1
        true_mean, _, true_log_var = ns.q_posterior(x_start, x_noisy, times)
        frozen = torch.cat([pred.detach(), out.chunk(2, dim=1)[1]], dim=1)
        mean, _, log_var = self._mean_variance_from_output(frozen, x_noisy, times, ns, True, predict_x_start)
        kl = meanflat(normal_kl(true_mean, true_log_var, mean, log_var)) * NAT
        nll = -meanflat(discretized_gaussian_log_likelihood(x_start, means=mean, log_scales=0.5 * log_var)) * NAT
        vb = torch.where(times == 0, nll, kl).mean()
        return loss + vb * self.vb_loss_weight

    def forward(self, image, text=None, image_embed=None, text_encodings=None, text_mask=None, unet_number=1):
        i = unet_number - 1
        unet, ns, size = self.unets[i], self.noise_schedulers[i], self.image_sizes[i]
        b = image.shape[0]
        times = ns.sample_random_times(b, image.device)
        if image_embed is None:
            image_embed, _ = self.clip.embed_image(image)
        if text is not None and text_encodings is None:
            _, text_encodings, text_mask = self.clip.embed_text(text)
        lowres = None
        if unet.lowres_cond:
            lowres = normalize_neg_one_to_one(self.to_lowres_cond(image, target_image_size=size,
                                                                  downsample_image_size=self.image_sizes[i - 1]))
        x = normalize_neg_one_to_one(resize_image_to(image, size))
        return self.p_losses(unet, x, times, ns, image_embed=image_embed,
                             text_encodings=text_encodings if unet.cond_on_text_encodings else None,
                             text_mask=text_mask if unet.cond_on_text_encodings else None, lowres_cond_img=lowres,
                             learned_variance=self.learned_variance[i], predict_x_start=self.predict_x_start[i])
