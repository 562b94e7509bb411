"""Loss trajectory of a small GPT-2 under a given parallel layout (run directly for 1 GPU, under torchrun for more).

    python tests/gpu_tp_parity.py --out one.json
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/gpu_tp_parity.py --tp 2 --out tp2.json

``tests/test_gpu.py`` compares the trajectories: tensor parallel with sequence parallelism and the collectives fused
into the GEMM kernels (CUDA-graph replayed blocks) must train like the single-GPU native path (same data, same init:
parameter initialisation is layout independent)."""
import argparse
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tp", type=int, default=1)
    ap.add_argument("--pp", type=int, default=1)
    ap.add_argument("--fused", type=int, default=1)
    ap.add_argument("--graphs", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--dropout", type=float, default=0.0)
    ap.add_argument("--zero", type=int, default=0)
    ap.add_argument("--out", default="")
    a = ap.parse_args()

    from libai_b200.config import LazyConfig
    from libai_b200.engine import DefaultTrainer, default_setup
    from libai_b200.utils import distributed as dutil

    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    cfg = LazyConfig.load(os.path.join(REPO, "configs", "gpt2_synthetic.py"))
    m = cfg.model.cfg
    m.hidden_layers, m.hidden_size, m.num_attention_heads, m.ffn_hidden_size = 4, 512, 8, 2048
    m.max_seq_length, m.vocab_size = 256, 4096
    m.embedding_dropout_prob = m.attention_dropout_prob = m.output_dropout_prob = a.dropout
    for ds in cfg.dataloader.train.dataset:
        ds.seq_length, ds.vocab_size = 256, 4096
    cfg.dataloader.train.num_workers = 0
    dp = world // (a.tp * a.pp)
    acc = 1 if a.pp == 1 else 4
    cfg.train.train_micro_batch_size = 16 // dp // acc           # global batch 16 x 256 tokens in every layout
    cfg.train.num_accumulation_steps = acc
    cfg.train.global_batch_size = None
    cfg.train.train_iter, cfg.train.log_period = 10 ** 6, 1
    cfg.train.amp.enabled = True
    cfg.train.evaluation.enabled = False
    cfg.train.checkpointer.period = 10 ** 9
    cfg.train.output_dir = os.path.join(REPO, "output", "tp_parity")
    cfg.train.cuda_graphs.enabled = bool(a.graphs)
    cfg.train.dist.tensor_parallel_size, cfg.train.dist.pipeline_parallel_size = a.tp, a.pp
    cfg.train.dist.pipeline_num_layers = 4
    cfg.train.dist.data_parallel_size = dp
    cfg.train.dist.sequence_parallel = bool(a.fused) and a.tp > 1
    cfg.train.dist.fused_tp_comm = bool(a.fused) and a.tp > 1
    cfg.train.zero_optimization.enabled = a.zero > 0
    cfg.train.zero_optimization.stage = max(a.zero, 1)
    cfg.optim.lr = 4e-4
    cfg.train.warmup_ratio = 0.0
    default_setup(cfg, argparse.Namespace(resume=False, config_file=""))
    trainer = DefaultTrainer(cfg)
    step = trainer._trainer
    # identical data in every layout: the full global batch is drawn from one generator and sliced per dp rank
    g = torch.Generator().manual_seed(77)
    topo = dutil.get_dist_util()
    losses = []
    for i in range(a.steps):
        toks = torch.randint(0, 4096, (16, 257), generator=g)
        toks[:, 1::2] = toks[:, 0:-1:2]          # learnable structure: every odd token repeats its predecessor
        mine = toks[topo.dp_rank * (16 // dp): (topo.dp_rank + 1) * (16 // dp)].cuda()
        mb = mine.shape[0] // acc
        batches = [dict(input_ids=mine[k * mb:(k + 1) * mb, :-1].contiguous(), labels=mine[k * mb:(k + 1) * mb, 1:].contiguous())
                   for k in range(acc)]
        out = step.train_on_batches(batches)
        lv = torch.zeros(1, device="cuda")
        if out:
            lv += sum(v for k, v in out.items() if "loss" in k).float()
        if world > 1:
            # mean over dp ranks; pipeline: only the last stage holds the loss
            torch.distributed.all_reduce(lv)
            holders = dp * a.tp
            lv /= holders
        losses.append(float(lv))
    from libai_b200 import ops

    res = {"losses": losses, "graphs": bool(step.graphs_enabled), "layout": str(topo), "launches": ops.launch_count()}
    if dutil.get_rank() == 0:
        print(json.dumps(res))
        if a.out:
            with open(a.out, "w") as f:
                json.dump(res, f)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
