"""Hierarchical volume rendering around two NeRF MLPs.

Spec: reference projects/NeRF/modeling/System.py:25-463 — coarse stratified sampling along each ray, alpha
compositing, inverse-CDF importance sampling for the fine network, and the train/eval/render output conventions of
``forward``.  ``rays``: [N, 8] = origin(3) direction(3) near(1) far(1).
"""
import collections

import torch
from torch import nn

from libai_b200.config import configurable, instantiate

from .nerf import Embedding, NeRF


def sample_pdf(bins, weights, n_samples, det=False, eps=1e-5):
    """Draw ``n_samples`` depths per ray from the piecewise-constant pdf ``weights`` over ``bins``
    ([N, S+1] edges, [N, S] weights)."""
    weights = weights + eps
    pdf = weights / weights.sum(-1, keepdim=True)
    cdf = torch.cat([torch.zeros_like(pdf[:, :1]), pdf.cumsum(-1)], dim=-1)        # [N, S+1]
    if det:
        u = torch.linspace(0, 1, n_samples, device=bins.device, dtype=bins.dtype).expand(bins.shape[0], n_samples)
    else:
        u = torch.rand(bins.shape[0], n_samples, device=bins.device, dtype=bins.dtype)
    u = u.contiguous()
    idx = torch.searchsorted(cdf.detach(), u, right=True)
    below = (idx - 1).clamp(min=0)
    above = idx.clamp(max=cdf.shape[-1] - 1)
    cdf_lo, cdf_hi = cdf.gather(1, below), cdf.gather(1, above)
    bin_lo, bin_hi = bins.gather(1, below), bins.gather(1, above)
    denom = cdf_hi - cdf_lo
    denom = torch.where(denom < eps, torch.ones_like(denom), denom)   # zero-probability bin → take its left edge
    return bin_lo + (u - cdf_lo) / denom * (bin_hi - bin_lo)


class NerfSystem(nn.Module):
    @configurable
    def __init__(self, D=8, W=256, in_channels_xyz=63, in_channels_dir=27, skips=(4,), N_samples=64, use_disp=False,
                 perturb=1.0, noise_std=1.0, N_importance=128, chunk=32 * 1024, dataset_type="Blender",
                 loss_func=None):
        super().__init__()
        self.N_samples, self.use_disp, self.perturb = N_samples, use_disp, perturb
        self.noise_std, self.N_importance, self.chunk = noise_std, N_importance, chunk
        self.white_back = dataset_type == "Blender"
        self.loss_func = nn.MSELoss() if loss_func is None else loss_func
        self.embedding_xyz = Embedding(3, (in_channels_xyz // 3 - 1) // 2)
        self.embedding_dir = Embedding(3, (in_channels_dir // 3 - 1) // 2)
        self.nerf_coarse = NeRF(D=D, W=W, input_ch=in_channels_xyz, input_ch_views=in_channels_dir, skips=skips)
        if N_importance > 0:
            self.nerf_fine = NeRF(D=D, W=W, input_ch=in_channels_xyz, input_ch_views=in_channels_dir, skips=skips)

    @classmethod
    def from_config(cls, cfg):
        keys = ("D", "W", "in_channels_xyz", "in_channels_dir", "skips", "N_samples", "use_disp", "perturb",
                "noise_std", "N_importance", "chunk", "dataset_type", "loss_func")
        out = {k: cfg[k] for k in keys if k in cfg}
        if hasattr(out.get("loss_func"), "keys") and "_target_" in out["loss_func"]:
            out["loss_func"] = instantiate(out["loss_func"])
        return out

    # ------------------------------------------------------------------ rendering
    def _composite(self, model, xyz, dirs_embedded, z_vals, ray_norm, weights_only=False):
        """Run ``model`` on the sample points and alpha-composite along each ray."""
        n_rays, n_samp = xyz.shape[:2]
        pts = self.embedding_xyz(xyz.reshape(-1, 3))
        dtype = next(model.parameters()).dtype
        out_chunks = []
        for i in range(0, pts.shape[0], self.chunk):
            p = pts[i:i + self.chunk]
            if weights_only:
                out_chunks.append(model(p.to(dtype), sigma_only=True))
            else:
                d = dirs_embedded.repeat_interleave(n_samp, dim=0)[i:i + self.chunk]
                out_chunks.append(model(torch.cat([p, d], dim=-1).to(dtype)))
        out = torch.cat(out_chunks, dim=0).float()
        if weights_only:
            sigmas = out.view(n_rays, n_samp)
        else:
            out = out.view(n_rays, n_samp, 4)
            rgbs, sigmas = out[..., :3], out[..., 3]
        deltas = torch.cat([z_vals[:, 1:] - z_vals[:, :-1], torch.full_like(z_vals[:, :1], 1e10)], dim=-1) * ray_norm
        noise = torch.randn_like(sigmas) * self.noise_std if (self.training and self.noise_std > 0) else 0.0
        alphas = 1.0 - torch.exp(-deltas * torch.relu(sigmas + noise))
        trans = torch.cumprod(torch.cat([torch.ones_like(alphas[:, :1]), 1.0 - alphas + 1e-10], dim=-1), dim=-1)[:, :-1]
        weights = alphas * trans
        if weights_only:
            return weights
        acc = weights.sum(-1)
        rgb = (weights[..., None] * rgbs).sum(-2)
        depth = (weights * z_vals).sum(-1)
        if self.white_back:
            rgb = rgb + (1.0 - acc[:, None])
        return rgb, depth, weights, acc

    def render_rays(self, rays, test_time=False):
        rays = rays.float()
        rays_o, rays_d, near, far = rays[:, 0:3], rays[:, 3:6], rays[:, 6:7], rays[:, 7:8]
        n_rays = rays.shape[0]
        dirs_embedded = self.embedding_dir(rays_d)
        ray_norm = rays_d.norm(dim=-1, keepdim=True)

        t = torch.linspace(0, 1, self.N_samples, device=rays.device)
        z_vals = (1.0 / (1.0 / near * (1 - t) + 1.0 / far * t)) if self.use_disp else near * (1 - t) + far * t
        z_vals = z_vals.expand(n_rays, self.N_samples)
        if self.perturb > 0 and self.training:
            mid = 0.5 * (z_vals[:, :-1] + z_vals[:, 1:])
            upper, lower = torch.cat([mid, z_vals[:, -1:]], -1), torch.cat([z_vals[:, :1], mid], -1)
            z_vals = lower + (upper - lower) * self.perturb * torch.rand_like(z_vals)
        xyz = rays_o[:, None] + rays_d[:, None] * z_vals[..., None]

        result = {}
        if test_time and self.N_importance > 0:
            weights_coarse = self._composite(self.nerf_coarse, xyz, dirs_embedded, z_vals, ray_norm, weights_only=True)
            result["opacity_coarse"] = weights_coarse.sum(-1)
        else:
            rgb, depth, weights_coarse, acc = self._composite(self.nerf_coarse, xyz, dirs_embedded, z_vals, ray_norm)
            result.update(rgb_coarse=rgb, depth_coarse=depth, opacity_coarse=acc)
        if self.N_importance > 0:
            mid = 0.5 * (z_vals[:, :-1] + z_vals[:, 1:])
            z_fine = sample_pdf(mid, weights_coarse[:, 1:-1].detach(), self.N_importance,
                                det=(self.perturb == 0 or not self.training)).detach()
            z_all = torch.sort(torch.cat([z_vals, z_fine], dim=-1), dim=-1).values
            xyz = rays_o[:, None] + rays_d[:, None] * z_all[..., None]
            rgb, depth, _, acc = self._composite(self.nerf_fine, xyz, dirs_embedded, z_all, ray_norm)
            result.update(rgb_fine=rgb, depth_fine=depth, opacity_fine=acc)
        return result

    def forward_features(self, rays):
        """Chunked ``render_rays`` (chunking matters for whole-image evaluation)."""
        results = collections.defaultdict(list)
        for i in range(0, rays.shape[0], self.chunk):
            for k, v in self.render_rays(rays[i:i + self.chunk]).items():
                results[k].append(v)
        return {k: torch.cat(v, dim=0) for k, v in results.items()}

    def _loss(self, results, rgbs):
        loss = self.loss_func(results["rgb_coarse"], rgbs.float())
        if "rgb_fine" in results:
            loss = loss + self.loss_func(results["rgb_fine"], rgbs.float())
        return loss

    def forward(self, rays, rgbs=None, c2w=None, valid_mask=None):
        """Training (``c2w is None``): ``{"losses"}``.  Validation (``c2w`` and ``rgbs``): losses + the rendered
        maps + ground truth.  Rendering a novel pose (``rgbs is None``): the rendered maps only."""
        rays = rays.reshape(-1, rays.shape[-1])
        if rgbs is not None:
            rgbs = rgbs.reshape(-1, 3)
        results = self.forward_features(rays)
        if c2w is None:
            return {"losses": self._loss(results, rgbs)}
        typ = "fine" if "rgb_fine" in results else "coarse"
        out = collections.OrderedDict()
        if rgbs is not None:
            out["losses"] = self._loss(results, rgbs)
        out[typ] = torch.zeros(1, device=rays.device)
        for k, v in results.items():
            out[k] = v.unsqueeze(0)
        if rgbs is not None:
            out["rgbs"] = rgbs.unsqueeze(0)
        return out
