"""Tile-N / split-K sweep of the wgrad (TN) GEMM on the benchmark model's shapes (dW[N_out, K_in] = dYᵀ·X over 8192
tokens).  Timings include the zero-fill of the fp32 output for the split-K variants (training accumulates in place)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))
from libai_b200.ops import load_ext  # noqa: E402


def timeit(fn, iters=30, warmup=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ext = load_ext()
    T = 8192
    out = {}
    for M, N in ((4096, 1024), (1024, 4096), (3072, 1024), (1024, 1024)):
        a = torch.randn(T, M, device="cuda").bfloat16()
        b = torch.randn(T, N, device="cuda").bfloat16()
        zero = torch.zeros(M, N, device="cuda")
        t_zero = timeit(lambda: zero.zero_())
        res = {"zero_fill_ms": round(t_zero, 4)}
        for bn in (0, 128, 192, 256):
            for s in (0, 1, 2, 3, 4, 6, 8):
                try:
                    ms = timeit(lambda: ext.gemm_tuned(a, b, 2, bn, s, True))
                except Exception as e:  # noqa
                    res[f"bn{bn}_s{s}"] = str(e)[:40]
                    continue
                res[f"bn{bn}_s{s}"] = round(ms, 4)
        res["cublas_ms"] = round(timeit(lambda: a.t() @ b), 4)
        out[f"{M}x{N}x{T}"] = res
        print(M, N, json.dumps(res), flush=True)
    json.dump(out, open("gpurun_out/wgrad_sweep.json", "w"), indent=1)


if __name__ == "__main__":
    main()
