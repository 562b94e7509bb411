import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from libai_b200 import ops
ext = ops.load_ext()
def rel(a, b): return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-9)).item()
for (M, K, N) in [(1024, 512, 2048), (8192, 1024, 4096), (256, 256, 512), (1024, 1024, 2048), (1024, 512, 1024)]:
    g = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(K, N, device="cuda") * 0.05).bfloat16()
    pre = torch.randn(M, N, device="cuda").bfloat16()
    try:
        out = ext.dgrad_actgrad(g, w, pre, 1)
        pf = pre.float().requires_grad_(True)
        torch.nn.functional.gelu(pf).backward(g.float() @ w.float())
        print((M, K, N), "ok rel", rel(out, pf.grad))
    except Exception as e:
        print((M, K, N), "FAIL", str(e)[:120])
    try:
        ext.gemm(g, w, 1, None, None, False, torch.bfloat16)
        print((M, K, N), "plain dgrad ok")
    except Exception as e:
        print((M, K, N), "plain FAIL", str(e)[:120])
