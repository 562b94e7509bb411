"""fp8 forward path: host-side plumbing (switches, cache invalidation, config key).  Numerics of the E4M3 GEMM and the
quantiser are checked on the GPU (tests/gpu_kernel_check.py "fp8 forward GEMM", tests/test_gpu.py)."""
import torch

from libai_b200 import ops
from libai_b200.config import LazyConfig
from libai_b200.ops import functional as OF


def test_switch_and_epoch():
    assert not ops.fp8_enabled()
    ops.set_fp8(True)
    try:
        assert ops.fp8_enabled()
        e = ops.fp8_weight_epoch()
        ops.bump_fp8_weight_epoch()
        assert ops.fp8_weight_epoch() == e + 1
        # CPU tensors never take the native path, fp8 or not
        x, w = torch.randn(4, 32), torch.randn(16, 32)
        assert torch.allclose(OF.linear(x, w), x @ w.t(), atol=1e-5)
    finally:
        ops.set_fp8(False)


def test_fp8_gate_needs_k_multiple_of_16():
    ops.set_fp8(True)
    try:
        assert OF._fp8_ok(torch.empty(8, 64), torch.empty(16, 64))
        assert not OF._fp8_ok(torch.empty(8, 72), torch.empty(16, 72))          # 72 % 16 != 0 -> bf16 kernel
        assert not OF._fp8_ok(torch.empty(8, 64), torch.empty(64, 16).t())      # non-contiguous weight
    finally:
        ops.set_fp8(False)
    assert not OF._fp8_ok(torch.empty(8, 64), torch.empty(16, 64))


def test_optimizer_step_invalidates_cached_weights():
    from libai_b200.optim import AdamW

    p = torch.nn.Parameter(torch.randn(8, 8))
    opt = AdamW([{"params": [p]}], lr=1e-3)
    p.grad = torch.ones_like(p)
    e = ops.fp8_weight_epoch()
    opt.step()
    assert ops.fp8_weight_epoch() > e


def test_config_key():
    cfg = LazyConfig.load("configs/gpt2_pretrain.py")
    assert cfg.train.fp8.enabled is False
    cfg = LazyConfig.apply_overrides(cfg, ["train.fp8.enabled=true"])
    assert cfg.train.fp8.enabled is True
