"""Stress the fused ZeRO kernels under artificial rank skew and localise a mismatch: after every fused reduce-scatter the
owned gradient slice is compared with NCCL's reduce-scatter of the same inputs, after every fused Adam + all-gather
the full bf16 parameter buffer is compared across ranks.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/gpu_zero_stress.py [iters]
"""
import json
import os
import random
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    from libai_b200.config import DictConfig
    from libai_b200.optim import AdamW
    from libai_b200.utils import distributed as dutil

    topo = dutil.setup_dist_util(DictConfig(dict(data_parallel_size=world, tensor_parallel_size=1, pipeline_parallel_size=1)))
    torch.manual_seed(5)
    params = [torch.nn.Parameter((torch.randn(1 << 22, device="cuda") * 0.02).bfloat16()),
              torch.nn.Parameter((torch.randn(1000, 333, device="cuda") * 0.02).bfloat16())]
    opt = AdamW(params, lr=1e-3, weight_decay=0.01)
    opt.fused_zero_comm = True
    opt.configure(zero_stage=1)
    opt.setup()
    fg = opt._groups[0]
    assert fg.symm is not None, "fused path not active"
    rng = random.Random(1234 + rank)
    rs_bad, ag_bad, rs_worst, first_bad = 0, 0, 0.0, None
    for it in range(iters):
        opt.zero_grad()
        gen = torch.Generator(device="cuda").manual_seed(100 * it + rank)
        local = torch.randn(fg.numel, device="cuda", generator=gen)
        if rng.random() < 0.5:
            torch.cuda._sleep(rng.randint(0, 3_000_000))       # up to ~1.5 ms of skew before the gradients exist
        fg.grad_flat.copy_(local)
        if rng.random() < 0.5:
            torch.cuda._sleep(rng.randint(0, 3_000_000))
        opt.sync_gradients()                                   # fused reduce-scatter
        ref = torch.empty(fg.hi - fg.lo, device="cuda")
        dist.reduce_scatter_tensor(ref, local / world, group=topo.dp_group)
        err = float((fg.grad_shard() - ref).abs().max())
        rs_worst = max(rs_worst, err)
        if err > 1e-4:
            rs_bad += 1
            first_bad = first_bad or ("rs", it, err)
        if rng.random() < 0.5:
            torch.cuda._sleep(rng.randint(0, 3_000_000))
        opt.step()                                             # fused Adam + all-gather
        mine = fg.param_flat.clone()
        dist.broadcast(mine, src=0, group=topo.dp_group)
        diff = float((fg.param_flat.float() - mine.float()).abs().max())
        if diff != 0.0:
            ag_bad += 1
            first_bad = first_bad or ("ag", it, diff)
    out = dict(rank=rank, world=world, iters=iters, rs_bad=rs_bad, ag_bad=ag_bad, rs_worst_abs_err=rs_worst, first_bad=first_bad)
    gathered = [None] * world
    dist.all_gather_object(gathered, out)
    if rank == 0:
        print(json.dumps(dict(ok=all(g["rs_bad"] == 0 and g["ag_bad"] == 0 for g in gathered), per_rank=gathered)))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
