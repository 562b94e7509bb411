"""One GEMM launch for ncu: python dev/ncu_gemm.py M N K layout"""
import sys

import torch

sys.path.insert(0, ".")
from libai_b200.ops import load_ext

M, N, K, layout = (int(x) for x in sys.argv[1:5])
ext = load_ext()
a = torch.randn(M, K, device="cuda").bfloat16()
b = torch.randn(N, K, device="cuda").bfloat16() if layout == 0 else torch.randn(K, N, device="cuda").bfloat16()
for _ in range(3):
    y = ext.gemm(a, b, layout, None, None, False, torch.bfloat16)
torch.cuda.synchronize()
print(float(y.float().abs().mean()))
