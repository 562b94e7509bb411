"""MoCo v3: momentum-contrast with a ViT backbone, projector + predictor MLPs, symmetric InfoNCE.

Spec: reference projects/MOCOV3/modeling/moco.py — ``_build_mlp`` with BatchNorm (:66-84), momentum update
(:90-95), ``contrastive_loss`` against the keys of *all* ranks with rank-offset labels, scaled by ``2T`` (:97-114),
cosine momentum schedule (:116-119), two-crop forward (:121-143), ``MoCo_ViT`` (:146-160).
The momentum schedule is driven by an internal step counter, so the model plugs into the default trainer.
"""
import copy
import math

import torch
import torch.distributed as dist
from torch import nn

from libai_b200.config import configurable
from libai_b200.utils import distributed as dutil


class MoCo(nn.Module):
    @configurable
    def __init__(self, base_encoder, momentum_encoder=None, dim=256, mlp_dim=4096, T=1.0, m=0.99, max_iter=300):
        super().__init__()
        self.T, self.m, self.max_iter = T, m, max_iter
        self.base_encoder = base_encoder
        self.momentum_encoder = momentum_encoder if momentum_encoder is not None else copy.deepcopy(base_encoder)
        self.base_encoder.num_classes = self.momentum_encoder.num_classes = dim
        self._build_projector_and_predictor_mlps(dim, mlp_dim)
        for pb, pm in zip(self.base_encoder.parameters(), self.momentum_encoder.parameters()):
            pm.data.copy_(pb.data)
            pm.requires_grad = False
        self.register_buffer("cu_iter", torch.zeros((), dtype=torch.long))

    @classmethod
    def from_config(cls, cfg):
        return {k: cfg[k] for k in ("base_encoder", "momentum_encoder", "dim", "mlp_dim", "T", "m", "max_iter") if k in cfg}

    @staticmethod
    def _build_mlp(num_layers, input_dim, mlp_dim, output_dim, last_bn=True):
        layers = []
        for i in range(num_layers):
            d1 = input_dim if i == 0 else mlp_dim
            d2 = output_dim if i == num_layers - 1 else mlp_dim
            layers.append(nn.Linear(d1, d2, bias=False))
            if i < num_layers - 1:
                layers += [nn.BatchNorm1d(d2), nn.ReLU(inplace=True)]
            elif last_bn:
                layers.append(nn.BatchNorm1d(d2, affine=False))
        return nn.Sequential(*layers)

    def _build_projector_and_predictor_mlps(self, dim, mlp_dim):
        raise NotImplementedError

    @torch.no_grad()
    def _update_momentum_encoder(self, m):
        for pb, pm in zip(self.base_encoder.parameters(), self.momentum_encoder.parameters()):
            pm.data.mul_(m).add_(pb.data.to(pm.dtype), alpha=1.0 - m)

    def contrastive_loss(self, q, k):
        q, k = nn.functional.normalize(q.float(), dim=1), nn.functional.normalize(k.float(), dim=1)
        topo = dutil.get_dist_util()
        rank, world = 0, 1
        if dist.is_available() and dist.is_initialized() and topo.data_parallel_size > 1:
            world, rank = topo.data_parallel_size, topo.dp_rank
            parts = [torch.empty_like(k) for _ in range(world)]
            dist.all_gather(parts, k.contiguous(), group=topo.dp_group)
            k = torch.cat(parts, dim=0)
        logits = torch.einsum("nc,mc->nm", q, k) / self.T
        labels = torch.arange(q.shape[0], dtype=torch.long, device=q.device) + q.shape[0] * rank
        return nn.functional.cross_entropy(logits, labels) * (2 * self.T)

    def adjust_moco_momentum(self, cu_iter, m):
        return 1.0 - 0.5 * (1.0 + math.cos(math.pi * cu_iter / self.max_iter)) * (1.0 - m)

    def forward(self, images, labels=None, cu_iter=None, m=None):
        if not self.training:
            return self.base_encoder(images)
        x1, x2 = torch.chunk(images, 2, dim=1)  # the two crops are stacked on the channel axis
        q1 = self.predictor(self.base_encoder(x1)["prediction_scores"].float())
        q2 = self.predictor(self.base_encoder(x2)["prediction_scores"].float())
        step = int(self.cu_iter) if cu_iter is None else cu_iter
        mom = self.adjust_moco_momentum(step, self.m if m is None else m)
        with torch.no_grad():
            self._update_momentum_encoder(mom)
            k1 = self.momentum_encoder(x1)["prediction_scores"]
            k2 = self.momentum_encoder(x2)["prediction_scores"]
            self.cu_iter += 1
        return {"losses": self.contrastive_loss(q1, k2) + self.contrastive_loss(q2, k1)}


class MoCo_ViT(MoCo):
    def _build_projector_and_predictor_mlps(self, dim, mlp_dim):
        hidden = self.base_encoder.head.weight.shape[1]
        self.base_encoder.head = self._build_mlp(3, hidden, mlp_dim, dim)
        self.momentum_encoder.head = self._build_mlp(3, hidden, mlp_dim, dim)
        self.predictor = self._build_mlp(2, dim, mlp_dim, dim)
