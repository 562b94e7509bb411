#!/usr/bin/env python
"""Headline benchmark: GPT-2 pre-training throughput (tokens/s) on N B200 GPUs of one node.

Model/config = the reference's published GPT-2 benchmark (docs Benchmark.md:20-26, BASELINE.md rows
6-8: 24 layers, hidden 1024, 16 heads, seq 1024, data parallel, micro-batch per GPU fixed → weak
scaling), bf16 compute with fp32 master weights, synthetic tokens, random-init weights.

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29500 bench.py --gpus 8 --steps 20 --warmup 5
    python bench.py --impl reference ...      # the unmodified reference (needs OneFlow)

Prints ONE JSON line (rank 0).  ``value`` is device-timed (CUDA events, max over ranks) over exactly
``--steps`` full training steps (fwd + bwd + grad sync + optimizer); ``e2e`` repeats the
measurement through the public trainer API including the per-step pinned-host→device input copy and
the device→host loss read.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

# published LiBai numbers for the GPT-2 data-parallel rows (samples/s × 1024 tokens), BASELINE.md #6-#8
PUBLISHED_TOKENS_PER_S = {1: 17940.0, 4: 64973.0, 8: 128655.0}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "ref", "reference"])
    ap.add_argument("--micro-batch", type=int, default=8, help="samples per GPU per step")
    ap.add_argument("--layers", type=int, default=24)
    ap.add_argument("--hidden", type=int, default=1024)
    ap.add_argument("--heads", type=int, default=16)
    ap.add_argument("--seq", type=int, default=1024)
    ap.add_argument("--tp", type=int, default=1)
    ap.add_argument("--pp", type=int, default=1)
    ap.add_argument("--zero", type=int, default=-1, help="ZeRO stage; -1 = auto (stage 1 with the fused NVLink kernels when dp > 1)")
    ap.add_argument("--acc", type=int, default=1)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--graphs", type=int, default=1,
                    help="1 (default): replay the transformer blocks from CUDA graphs (data-parallel layouts only; "
                         "tensor/pipeline-parallel runs fall back to eager launches)")
    ap.add_argument("--fused-bias-grad", type=int, default=-1,
                    help="1/0: bias gradient of the MLP's first linear inside the dgrad epilogue (-1: library default)")
    ap.add_argument("--fp8", type=int, default=0,
                    help="1: forward GEMMs with E4M3 operands (experiment; the headline number is the bf16 default — "
                         "the JSON line then says dtype fp8-fwd/bf16-bwd)")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------
# reference arm
# ---------------------------------------------------------------------------------------------------
def run_reference(args):
    """Run the UNMODIFIED reference (baseline/_ref) through its own public API.  LiBai is a thin
    layer over OneFlow; without OneFlow (not in the image, no network) it cannot execute."""
    ref_dir = os.path.join(REPO, "baseline", "_ref")
    why = None
    if not os.path.isdir(ref_dir):
        why = "baseline/_ref not installed"
    else:
        sys.path.insert(0, ref_dir)
        try:
            import oneflow  # noqa: F401
        except Exception as e:  # noqa
            why = f"reference needs OneFlow which is not installed/installable offline ({type(e).__name__}: {e})"
    if why is None:
        try:
            import libai  # noqa: F401

            why = "reference import succeeded but no OneFlow CUDA runtime path is wired for sm_100"
        except Exception as e:  # noqa
            why = f"import libai failed: {type(e).__name__}: {e}"
    print(json.dumps({"impl": "reference", "unavailable": why.replace("\n", " ")[:300]}))
    return 0


# ---------------------------------------------------------------------------------------------------
# clocks sampler
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu_index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return

        def pump():
            for line in self.proc.stdout:
                self.rows.append(line.strip())

        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for row in self.rows:
            parts = [p.strip() for p in row.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {
            "sm_mhz": statistics.median(sm) if sm else None,
            "sm_max_mhz": max(mx) if mx else None,
            "reasons": sorted(reasons),
            "samples": len(sm),
        }


# ---------------------------------------------------------------------------------------------------
# native / ref arms
# ---------------------------------------------------------------------------------------------------
def build_cfg(args, world):
    from libai_b200.config import LazyConfig

    if args.zero < 0:
        args.zero = 1 if world // (args.tp * args.pp) > 1 else 0
    cfg = LazyConfig.load(os.path.join(REPO, "configs", "gpt2_synthetic.py"))
    m = cfg.model.cfg
    m.hidden_layers, m.hidden_size, m.num_attention_heads = args.layers, args.hidden, args.heads
    m.ffn_hidden_size = 4 * args.hidden
    m.max_seq_length = args.seq
    for ds in cfg.dataloader.train.dataset:
        ds.seq_length = args.seq
        ds.vocab_size = m.vocab_size
    cfg.dataloader.train.num_workers = 2
    cfg.train.train_micro_batch_size = args.micro_batch
    cfg.train.num_accumulation_steps = args.acc
    cfg.train.global_batch_size = None
    cfg.train.train_iter = 10 ** 6
    cfg.train.log_period = 10 ** 9
    cfg.train.amp.enabled = True
    cfg.train.evaluation.enabled = False
    cfg.train.checkpointer.period = 10 ** 9
    cfg.train.output_dir = os.path.join(REPO, "output", "bench")
    cfg.train.dist.tensor_parallel_size = args.tp
    cfg.train.dist.pipeline_parallel_size = args.pp
    cfg.train.dist.pipeline_num_layers = args.layers
    cfg.train.dist.data_parallel_size = world // (args.tp * args.pp)
    cfg.train.zero_optimization.enabled = args.zero > 0
    cfg.train.zero_optimization.stage = max(args.zero, 1)
    return cfg


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.impl == "ref":
        os.environ["LIBAI_B200_IMPL"] = "ref"

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run"
    assert torch.cuda.is_available(), "bench.py needs a GPU"
    torch.cuda.set_device(local_rank)

    import logging

    from libai_b200 import ops
    from libai_b200.engine import DefaultTrainer, default_setup
    from libai_b200.utils import distributed as dutil

    cfg = build_cfg(args, world)
    default_setup(cfg, argparse.Namespace(resume=False, config_file=""))
    logging.getLogger("libai_b200").setLevel(logging.WARNING)
    torch.manual_seed(cfg.train.seed + rank)
    trainer = DefaultTrainer(cfg)  # public API: builds model, optimizer, scheduler, loader, hooks
    step = trainer._trainer
    topo = dutil.get_dist_util()
    dev = torch.device("cuda", local_rank)
    tokens_per_step = cfg.train.global_batch_size * args.seq

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-timed: batches staged on the device, CUDA events around exactly K steps ----------------
    staged = []
    it = iter(trainer.train_loader)
    for _ in range(args.acc * 4):
        staged.append(DefaultTrainer.get_batch(next(it), "cuda"))
    torch.cuda.synchronize()
    model, optimizer = trainer.model, trainer.optimizer
    acc = args.acc
    if args.fp8:
        from libai_b200 import ops as _ops

        _ops.set_fp8(True)
    if args.fused_bias_grad >= 0:
        from libai_b200 import ops as _ops

        _ops.set_fused_bias_grad(bool(args.fused_bias_grad))
    graphs_on = False
    if args.graphs and topo.pipeline_parallel_size == 1:
        from libai_b200.engine.cuda_graphs import enable_for_model

        graphs_on = enable_for_model(model, staged[0])
        step._graphs_tried, step.graphs_enabled = True, graphs_on
        optimizer.zero_grad()

    def one_step(i):
        optimizer.zero_grad()
        if topo.pipeline_parallel_size > 1:
            from libai_b200.parallel.pipeline import PipelineSchedule1F1B

            if step._pipeline is None:
                step._pipeline = PipelineSchedule1F1B(model)
            loss = step._pipeline.run([staged[(i * acc + k) % len(staged)] for k in range(acc)])
        else:
            loss = None
            for k in range(acc):
                step._arm_grad_overlap(k == acc - 1)   # same as StepTrainer.run_step: early DP reduce on the last micro-batch
                out = model(**staged[(i * acc + k) % len(staged)])
                l = sum(v for kk, v in out.items() if "loss" in kk) / acc
                l.backward()
                loss = l.detach()
        optimizer.step()
        return loss

    for i in range(args.warmup):
        one_step(i)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ops.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    t_enq = time.perf_counter()
    # `ncu --nvtx --nvtx-include "bench_step"` → exactly the timed steps.  A start/end range (process-wide), not
    # push/pop (per thread): the backward kernels are launched from the autograd engine's thread.
    nvtx = os.environ.get("LIBAI_B200_NVTX", "0") == "1"
    for i in range(args.steps):
        rid = torch.cuda.nvtx.range_start("bench_step") if nvtx else None
        loss = one_step(args.warmup + i)
        if nvtx:
            torch.cuda.synchronize()
            torch.cuda.nvtx.range_end(rid)
    e1.record()
    # host time spent ENQUEUEING the timed steps (no sync inside): close to the device time = the step is launch-bound
    enqueue_ms = max_over_ranks((time.perf_counter() - t_enq) * 1e3) / args.steps
    barrier()
    launches = ops.launch_count()
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    value = tokens_per_step / (dev_ms * 1e-3)

    # ---- end to end through the trainer API: loader (pinned host) → H2D → step → loss D2H ----------------
    e2e = None
    if not args.no_e2e:
        h2d = d2h = 0
        from libai_b200.utils.events import EventStorage

        with EventStorage(0) as storage:
            trainer.storage = storage
            step.log_period = 10 ** 9
            for i in range(max(2, args.warmup)):
                trainer.iter = i
                trainer.run_step()
            barrier()
            t0 = time.perf_counter()
            for i in range(args.steps):
                trainer.iter = 10 + i
                # run_step: next(loader) [pinned] → get_batch (H2D) → fwd/bwd/optimizer
                data_iter = step._data_loader_iter
                batches = []
                for _ in range(acc):
                    inst = next(data_iter)
                    h2d += sum(v.tensor.numel() * v.tensor.element_size() for v in inst.get_fields().values())
                    batches.append(DefaultTrainer.get_batch(inst, "cuda"))
                optimizer.zero_grad()
                if topo.pipeline_parallel_size > 1:
                    out = step._pipeline.run(batches)
                    lval = sum(v for kk, v in (out or {}).items() if "loss" in kk) if out else torch.zeros((), device=dev)
                else:
                    lval = None
                    for j, b in enumerate(batches):
                        step._arm_grad_overlap(j == len(batches) - 1)
                        out = model(**b)
                        l = sum(v for kk, v in out.items() if "loss" in kk) / acc
                        l.backward()
                        lval = l.detach() if lval is None else lval + l.detach()
                optimizer.step()
                host_loss = float(lval.float().cpu()) if torch.is_tensor(lval) else float(lval)  # D2H read (sync)
                d2h += 4
            barrier()
            wall = time.perf_counter() - t0
        wall = max_over_ranks(wall)
        e2e = {
            "value": tokens_per_step * args.steps / wall,
            "unit": "tokens/s",
            "h2d_bytes_per_step": h2d // args.steps,
            "d2h_bytes_per_step": d2h // args.steps,
            "last_loss": host_loss,
        }

    if rank == 0:
        par = f"dp{topo.data_parallel_size}"
        if topo.tensor_parallel_size > 1:
            par += f"_tp{topo.tensor_parallel_size}"
        if topo.pipeline_parallel_size > 1:
            par += f"_pp{topo.pipeline_parallel_size}"
        if args.zero:
            par += f"_zero{args.zero}"
        base = PUBLISHED_TOKENS_PER_S.get(world)
        line = {
            "metric": "tokens/sec GPT-2 (nl24 h1024 a16 s1024) pre-training, device-timed max-over-ranks",
            "value": value,
            "unit": "tokens/s",
            "n_gpus": world,
            "steps": args.steps,
            "warmup": args.warmup,
            "ms_per_step": dev_ms,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": (value / base) if base else None,
            "dtype": "bf16" if not args.fp8 else "fp8(e4m3)-fwd/bf16-bwd",
            "data": "synthetic tokens, random-init weights",
            "impl": args.impl,
            "config": {
                "model": f"GPT-2 nl{args.layers} h{args.hidden} a{args.heads} (reference benchmark model, ~355M params)",
                "global_batch": cfg.train.global_batch_size,
                "micro_batch_per_gpu": args.micro_batch,
                "seq_len": args.seq,
                "parallelism": par,
                "l2_policy": "working set (weights+activations+optimizer state >> 126MB L2) exceeds L2 every step",
                "cuda_graphs": graphs_on,
            },
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": launches,
            "host_enqueue_ms_per_step": enqueue_ms,
            "host_cpus": len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count(),
            "final_loss": float(loss) if loss is not None and torch.is_tensor(loss) else None,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
