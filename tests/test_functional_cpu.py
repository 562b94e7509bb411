"""Reference (CPU) paths of the fused functional ops agree with their unfused compositions."""
import torch

from libai_b200.ops import functional as OF


def test_linear_bias_residual_and_mlp_residual_reference_paths():
    torch.manual_seed(0)
    x, w, b, r = torch.randn(6, 16), torch.randn(8, 16), torch.randn(8), torch.randn(6, 8)
    assert torch.allclose(OF.linear_bias_residual(x, w, b, r), x @ w.t() + b + r, atol=1e-5)
    assert torch.allclose(OF.linear_bias_residual(x, w, None, None), x @ w.t(), atol=1e-5)
    w1, b1, w2, b2, res = torch.randn(12, 16), torch.randn(12), torch.randn(16, 12), torch.randn(16), torch.randn(6, 16)
    want = torch.nn.functional.gelu(x @ w1.t() + b1) @ w2.t() + b2 + res
    assert torch.allclose(OF.mlp(x, w1, b1, w2, "gelu", b2, res), want, atol=1e-4)
    assert torch.allclose(OF.mlp(x, w1, b1, w2, "gelu"), want - b2 - res, atol=1e-4)


def test_layer_norm_with_skip_reference_path():
    torch.manual_seed(0)
    x = torch.randn(4, 10, requires_grad=True)
    g, b = torch.randn(10), torch.randn(10)
    ln, skip = OF.layer_norm_with_skip(x, g, b)
    assert skip is x
    assert torch.allclose(ln, torch.nn.functional.layer_norm(x, (10,), g, b, 1e-5), atol=1e-5)
    (ln.sum() + (skip * 2).sum()).backward()
    assert x.grad is not None and torch.isfinite(x.grad).all()
