"""ops/gated_mlp.py: the zero-copy [2F, K] view over adjacent gate / up weights (and their main_grad slices)."""
import torch

from libai_b200.ops import gated_mlp


def test_fused_views_only_for_adjacent_parameters():
    f, k = 8, 16
    flat = torch.arange(3 * f * k, dtype=torch.float32)
    grad = torch.zeros(3 * f * k)
    wg, wu = flat[: f * k].view(f, k), flat[f * k : 2 * f * k].view(f, k)
    wg.main_grad, wu.main_grad = grad[: f * k].view(f, k), grad[f * k : 2 * f * k].view(f, k)
    w, mg = gated_mlp.fused_gate_up_views(wg, wu)
    assert w.shape == (2 * f, k) and torch.equal(w, torch.cat([wg, wu])) and w.data_ptr() == wg.data_ptr()
    mg += 1.0                                                     # writes through to both parameters' gradient slices
    assert float(wg.main_grad.sum()) == f * k and float(wu.main_grad.sum()) == f * k and float(grad[2 * f * k :].sum()) == 0
    # not adjacent (a gap, or separate allocations): no fused view
    far = flat[2 * f * k :].view(f, k)
    assert gated_mlp.fused_gate_up_views(wg, far) is None
    assert gated_mlp.fused_gate_up_views(torch.zeros(f, k), torch.zeros(f, k)) is None
    # adjacent weights without (adjacent) main_grad: weight view only
    a, b = flat[: f * k].view(f, k).clone(), None
    buf = torch.zeros(2 * f * k)
    a, b = buf[: f * k].view(f, k), buf[f * k :].view(f, k)
    w2, mg2 = gated_mlp.fused_gate_up_views(a, b)
    assert w2.shape == (2 * f, k) and mg2 is None


def test_llama_mlp_cpu_path_is_unchanged():
    from libai_b200.models.llama_model import LlamaMLP

    torch.manual_seed(0)
    mlp = LlamaMLP(32, 64)
    x = torch.randn(4, 32)
    ref = mlp.down_proj(torch.nn.functional.silu(mlp.gate_proj(x)) * mlp.up_proj(x))
    assert torch.allclose(mlp(x), ref, atol=1e-6)
