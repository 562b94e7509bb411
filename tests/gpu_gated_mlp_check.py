"""Gate/up projection as one GEMM (ops/gated_mlp.py) vs the two-GEMM path of LlamaMLP: outputs, input gradients and
main_grad accumulations must agree.  `python tests/gpu_gated_mlp_check.py` (1 GPU) or under torchrun with 2+ GPUs
(tensor parallel, fused collectives)."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def rel_err(a, b):
    return float((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-6))


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    from libai_b200.config import DictConfig
    from libai_b200.layers._param import param_defaults
    from libai_b200.models.llama_model import LlamaMLP
    from libai_b200.ops import gated_mlp
    from libai_b200.optim import AdamW
    from libai_b200.utils import distributed as dutil

    if world > 1:
        dutil.init_process_group("cuda")
    dutil.reset_dist_util()
    dutil.setup_dist_util(DictConfig(dict(data_parallel_size=1, tensor_parallel_size=world, pipeline_parallel_size=1,
                                          sequence_parallel=world > 1, fused_tp_comm=world > 1)))
    H, F, T = 1024, 2816, 4096
    with param_defaults(dtype=torch.bfloat16, device="cuda", seed=11):
        mlp = LlamaMLP(H, F)
    opt = AdamW([{"params": list(mlp.parameters())}], lr=1e-3)
    opt.setup()                                   # parameters -> flat buffer (gate / up adjacent), main_grad bound
    views = gated_mlp.fused_gate_up_views(mlp.gate_proj.weight, mlp.up_proj.weight)
    assert views is not None and views[1] is not None, "gate / up weights are not adjacent after optimizer.setup()"
    g = torch.Generator(device="cuda").manual_seed(5 + rank)
    x0 = (torch.randn(T // world, H, device="cuda", generator=g) * 0.5).bfloat16()
    gy = (torch.randn(T // world, H, device="cuda", generator=g) * 0.1).bfloat16()
    res = {}
    for mode in ("two_gemms", "fused"):
        gated_mlp.set_enabled(mode == "fused")
        opt.zero_grad()
        x = x0.clone().requires_grad_(True)
        y = mlp(x)
        y.backward(gy)
        torch.cuda.synchronize()
        res[mode] = dict(y=y.detach().clone(), gx=x.grad.clone(),
                         gg=mlp.gate_proj.weight.main_grad.clone(), gu=mlp.up_proj.weight.main_grad.clone(),
                         gd=mlp.down_proj.weight.main_grad.clone())
    errs = {k: rel_err(res["fused"][k], res["two_gemms"][k]) for k in res["fused"]}
    ok = max(errs.values()) < 2e-2
    # timing
    def timeit(fn, iters=20):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    times = {}
    for mode in ("two_gemms", "fused"):
        gated_mlp.set_enabled(mode == "fused")

        def step():
            x = x0.clone().requires_grad_(True)
            mlp(x).backward(gy)

        times[mode] = timeit(step)
    if world > 1:
        t = torch.tensor([times["two_gemms"], times["fused"], 0.0 if ok else 1.0], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times = {"two_gemms": float(t[0]), "fused": float(t[1])}
        ok = float(t[2]) == 0.0
    if rank == 0:
        print(json.dumps({"name": f"gated MLP fused gate/up vs two GEMMs (tp={world})", "ok": ok, "errs": errs,
                          "fwd_bwd_ms": times}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
