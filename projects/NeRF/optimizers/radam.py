"""Rectified Adam (Liu et al. 2019; reference projects/NeRF/optimizers/Radam.py).

The variance-rectification term only depends on the step count, so it is computed once per step on the host and the
parameter update runs as ``torch._foreach`` ops over the whole group — a handful of launches per step instead of a
handful per parameter."""
import math

import torch
from torch.optim import Optimizer


class RAdam(Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, degenerated_to_sgd=True):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1):
            raise ValueError("invalid RAdam hyper-parameters")
        self.degenerated_to_sgd = degenerated_to_sgd
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    @staticmethod
    def rectification(step, beta2):
        """(use_adaptive, step_scale): the length of the approximated SMA decides whether the second moment is
        trustworthy yet."""
        rho_inf = 2.0 / (1.0 - beta2) - 1.0
        beta2_t = beta2 ** step
        rho_t = rho_inf - 2.0 * step * beta2_t / (1.0 - beta2_t)
        if rho_t >= 5:
            r = math.sqrt((1 - beta2_t) * (rho_t - 4) / (rho_inf - 4) * (rho_t - 2) / rho_t * rho_inf / (rho_inf - 2))
            return True, r
        return False, 1.0

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            params = [p for p in group["params"] if p.grad is not None]
            if not params:
                continue
            beta1, beta2 = group["betas"]
            for p in params:
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, dtype=torch.float32)
                    st["exp_avg_sq"] = torch.zeros_like(p, dtype=torch.float32)
                st["step"] += 1
            step = self.state[params[0]]["step"]
            grads = [p.grad.float() for p in params]
            m = [self.state[p]["exp_avg"] for p in params]
            v = [self.state[p]["exp_avg_sq"] for p in params]
            torch._foreach_mul_(m, beta1)
            torch._foreach_add_(m, grads, alpha=1 - beta1)
            torch._foreach_mul_(v, beta2)
            torch._foreach_addcmul_(v, grads, grads, value=1 - beta2)
            adaptive, r = self.rectification(step, beta2)
            bias1 = 1 - beta1 ** step
            if not adaptive and not self.degenerated_to_sgd:
                continue
            master = [p.float() if p.dtype != torch.float32 else p for p in params]
            if group["weight_decay"] != 0:
                torch._foreach_mul_(master, 1 - group["weight_decay"] * group["lr"])
            if adaptive:
                denom = torch._foreach_sqrt(v)
                torch._foreach_add_(denom, group["eps"])
                torch._foreach_addcdiv_(master, m, denom, value=-group["lr"] * r / bias1)
            else:
                torch._foreach_add_(master, m, alpha=-group["lr"] / bias1)
            for p, w in zip(params, master):
                if w is not p:
                    p.copy_(w)
        return loss
