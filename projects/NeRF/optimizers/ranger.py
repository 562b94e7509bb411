"""Ranger = RAdam + Lookahead (Zhang et al. 2019), optionally with gradient centralisation (reference
projects/NeRF/optimizers/Ranger.py).  Every ``k`` steps the slow weights move ``alpha`` of the way to the fast
weights and the fast weights restart from there."""
import torch

from .radam import RAdam


class Ranger(RAdam):
    def __init__(self, params, lr=1e-3, alpha=0.5, k=6, N_sma_threshhold=5, betas=(0.95, 0.999), eps=1e-5,
                 weight_decay=0, use_gc=False):
        if not 0.0 <= alpha <= 1.0 or k < 1:
            raise ValueError("invalid Lookahead hyper-parameters")
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, degenerated_to_sgd=False)
        self.alpha, self.k, self.use_gc = alpha, k, use_gc
        for g in self.param_groups:
            g.setdefault("alpha", alpha)
            g.setdefault("k", k)

    @torch.no_grad()
    def step(self, closure=None):
        if self.use_gc:
            for group in self.param_groups:
                for p in group["params"]:
                    if p.grad is not None and p.grad.dim() > 1:
                        p.grad.sub_(p.grad.mean(dim=tuple(range(1, p.grad.dim())), keepdim=True))
        loss = super().step(closure)
        for group in self.param_groups:
            for p in group["params"]:
                st = self.state.get(p)
                if not st:
                    continue
                if "slow_buffer" not in st:
                    st["slow_buffer"] = p.detach().clone()
                if st["step"] % group["k"] == 0:
                    st["slow_buffer"].add_(p - st["slow_buffer"], alpha=group["alpha"])
                    p.copy_(st["slow_buffer"])
        return loss
