"""ZeRO stage 1 / 2 / 3 on two data-parallel CPU ranks (gloo): identical training trajectory, and the per-rank tensor
footprint drops with the stage (reference: graph_base.py:69-70 `enable_zero(True, stage)`; its tests run stage 3,
tests/models/test_gpt.py:185-199)."""
import os
import sys

import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "tools"))

TINY = [
    "model.cfg.hidden_layers=8", "model.cfg.hidden_size=64", "model.cfg.ffn_hidden_size=128",
    "model.cfg.num_attention_heads=4", "model.cfg.vocab_size=128", "model.cfg.max_seq_length=16",
    "dataloader.train.dataset.0.vocab_size=128", "dataloader.train.dataset.0.seq_length=16",
    "dataloader.train.dataset.0.num_samples=256", "dataloader.train.num_workers=0",
    "train.train_micro_batch_size=4", "train.log_period=1", "train.amp.enabled=false", "train.warmup_ratio=0.0",
    "train.dist.pipeline_num_layers=8", "train.dist.data_parallel_size=2", "train.evaluation.enabled=false",
    "optim.lr=1e-2", "train.train_iter=6", "train.checkpointer.period=3",
]


def _worker(rank, world, out_dir, stage, acc, resume):
    import json

    import train_net
    from libai_b200.config import default_argument_parser
    from libai_b200.engine import DefaultTrainer

    footprint = {}
    orig_train = DefaultTrainer.train

    def train(self):
        opt = self.optimizer
        storages = {}
        for fg in opt._groups:
            if fg is None:
                continue
            tensors = [getattr(fg, name, None) for name in ("param_flat", "grad_flat", "pool_buf", "pool_part", "master", "red", "param_shard")]
            tensors += list(fg.state.values())
            for t in tensors:
                if torch.is_tensor(t):
                    st = t.untyped_storage()
                    storages[st.data_ptr()] = st.nbytes()        # views share their storage: count it once
        persistent = sum(storages.values())
        footprint["bytes"] = persistent
        return orig_train(self)

    DefaultTrainer.train = train
    try:
        argv = ["--config-file", os.path.join(REPO, "configs/gpt2_synthetic.py")] + (["--resume"] if resume else [])
        extra = [f"train.output_dir={out_dir}", f"train.num_accumulation_steps={acc}",
                 f"train.zero_optimization.enabled={'true' if stage else 'false'}", f"train.zero_optimization.stage={max(stage, 1)}"]
        train_net.main(default_argument_parser().parse_args(argv + TINY + extra))
    finally:
        DefaultTrainer.train = orig_train
    losses = [json.loads(ln) for ln in open(os.path.join(out_dir, "metrics.json"))]
    return {"losses": [m["total_loss"] for m in losses if "total_loss" in m], "bytes": footprint["bytes"]}


@pytest.mark.parametrize("acc", [1, 2])
def test_zero_stages_match_and_shrink(tmp_path, acc):
    from tests.dist_utils import run_distributed

    res = {}
    for stage in (0, 1, 2, 3):
        out = run_distributed(_worker, 2, str(tmp_path / f"s{stage}_a{acc}"), stage, acc, False, timeout=600)
        res[stage] = out[0]
    base = res[0]["losses"]
    assert len(base) >= 6
    for stage in (1, 2, 3):
        for a, b in zip(base, res[stage]["losses"]):
            assert abs(a - b) < 2e-3 * max(1.0, abs(a)), (stage, base, res[stage]["losses"])
    # resident state per rank: stage 1 shards master + moments, stage 2 also the gradient buffers of the blocks, stage 3
    # also their parameters
    assert res[1]["bytes"] < res[0]["bytes"]
    assert res[2]["bytes"] < res[1]["bytes"]
    assert res[3]["bytes"] < res[2]["bytes"]
    # the final checkpoints (logical tensors) agree between stage 1 and stage 3
    a = torch.load(str(tmp_path / f"s1_a{acc}" / "model_final" / "model"), weights_only=False)
    b = torch.load(str(tmp_path / f"s3_a{acc}" / "model_final" / "model"), weights_only=False)
    assert a.keys() == b.keys()
    assert max((a[k].float() - b[k].float()).abs().max().item() for k in a) < 1e-4


def test_zero3_resume_is_exact(tmp_path):
    """Stage-3 checkpoint (written from gathered parameters, sharded optimizer state) → resume continues exactly."""
    import shutil

    from tests.dist_utils import run_distributed

    full, part = str(tmp_path / "full"), str(tmp_path / "part")
    run_distributed(_worker, 2, full, 3, 1, False, timeout=600)
    os.makedirs(part)
    shutil.copytree(os.path.join(full, "model_0000002"), os.path.join(part, "model_0000002"))
    with open(os.path.join(part, "last_checkpoint"), "w") as f:
        f.write("model_0000002")
    run_distributed(_worker, 2, part, 3, 1, True, timeout=600)
    a = torch.load(os.path.join(full, "model_final", "model"), weights_only=False)
    b = torch.load(os.path.join(part, "model_final", "model"), weights_only=False)
    assert a.keys() == b.keys() and max((a[k].float() - b[k].float()).abs().max().item() for k in a) < 1e-6


def _worker_tp(rank, world, out_dir, stage):
    import json

    import train_net
    from libai_b200.config import default_argument_parser

    argv = ["--config-file", os.path.join(REPO, "configs/gpt2_synthetic.py")]
    tiny = [t for t in TINY if not t.startswith("train.dist.data_parallel_size")]
    extra = [f"train.output_dir={out_dir}", "train.dist.tensor_parallel_size=2", "train.dist.data_parallel_size=2",
             "train.dist.sequence_parallel=true", "train.zero_optimization.enabled=true",
             f"train.zero_optimization.stage={stage}", "train.checkpointer.period=100"]
    train_net.main(default_argument_parser().parse_args(argv + tiny + extra))
    losses = [json.loads(ln) for ln in open(os.path.join(out_dir, "metrics.json"))]
    return [m["total_loss"] for m in losses if "total_loss" in m]


def test_zero2_and_3_with_tensor_and_sequence_parallelism(tmp_path):
    """tp2 x dp2 (4 gloo ranks): the per-block buckets carry the tensor-parallel fix-ups (all-reduce of the gradients of
    parameters replicated over TP but computed on token shards) before their data-parallel reduce-scatter — the
    trajectory must be the one of stage 1."""
    from tests.dist_utils import run_distributed

    res = {}
    for stage in (1, 2, 3):
        res[stage] = run_distributed(_worker_tp, 4, str(tmp_path / f"tp_s{stage}"), stage, timeout=900)[0]
    assert len(res[1]) >= 6
    for stage in (2, 3):
        for a, b in zip(res[1], res[stage]):
            assert abs(a - b) < 2e-3 * max(1.0, abs(a)), (stage, res[1], res[stage])


def _worker_pp(rank, world, out_dir, stage):
    import json

    import train_net
    from libai_b200.config import default_argument_parser

    argv = ["--config-file", os.path.join(REPO, "configs/gpt2_synthetic.py")]
    tiny = [t for t in TINY if not t.startswith("train.dist.data_parallel_size")]
    extra = [f"train.output_dir={out_dir}", "train.dist.pipeline_parallel_size=2", "train.dist.data_parallel_size=2",
             "train.num_accumulation_steps=2", "train.zero_optimization.enabled=true",
             f"train.zero_optimization.stage={stage}", "train.checkpointer.period=100"]
    train_net.main(default_argument_parser().parse_args(argv + tiny + extra))
    path = os.path.join(out_dir, "metrics.json")
    losses = [json.loads(ln) for ln in open(path)] if os.path.exists(path) else []
    return [m["total_loss"] for m in losses if "total_loss" in m]


def test_zero2_under_pipeline_parallelism(tmp_path):
    """pp2 x dp2, two 1F1B micro-batches: every micro-batch's backward opens, fills and reduce-scatters the pooled
    per-block gradient buffers again — same trajectory as stage 1."""
    from tests.dist_utils import run_distributed

    res = {}
    for stage in (1, 2):
        outs = run_distributed(_worker_pp, 4, str(tmp_path / f"pp_s{stage}"), stage, timeout=900)
        res[stage] = max(outs, key=len)          # the rank that logs the loss
    assert len(res[1]) >= 6
    for a, b in zip(res[1], res[2]):
        assert abs(a - b) < 2e-3 * max(1.0, abs(a)), (res[1], res[2])
