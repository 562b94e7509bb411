"""Distributed semantics on CPU/gloo: topology, TP layers vs nn.Linear, and loss/weight equality of a
tiny GPT across (dp, tp, pp, sequence-parallel, ZeRO) layouts against the single-process run."""
import pytest
import torch

from tests.dist_utils import run_distributed

TINY = dict(
    hidden_layers=4, vocab_size=128, hidden_size=32, ffn_hidden_size=64, num_attention_heads=4,
    max_seq_length=16, embedding_dropout_prob=0.0, attention_dropout_prob=0.0, output_dropout_prob=0.0,
    layernorm_epsilon=1e-5, initializer_range=0.02, use_scaled_init_for_output_weights=True,
    bias_gelu_fusion=True, bias_dropout_fusion=True, scale_mask_softmax_fusion=True,
    apply_query_key_layer_scaling=True, apply_residual_post_layernorm=False, amp_enabled=False,
)
GLOBAL_BATCH, N_STEPS = 8, 3


def _data(step):
    g = torch.Generator().manual_seed(100 + step)
    toks = torch.randint(0, TINY["vocab_size"], (GLOBAL_BATCH, TINY["max_seq_length"] + 1), generator=g)
    return toks[:, :-1].contiguous(), toks[:, 1:].contiguous()


def _train(rank, world, dp, tp, pp, sp, zero, acc):
    """Runs N_STEPS optimizer steps; returns (losses, {logical param name: tensor}) on rank 0."""
    from libai_b200.config import DictConfig
    from libai_b200.layers._param import param_defaults
    from libai_b200.models.gpt_model import GPTForPreTraining
    from libai_b200.optim import AdamW, get_default_optimizer_params
    from libai_b200.parallel import state as pstate
    from libai_b200.parallel.pipeline import PipelineSchedule1F1B
    from libai_b200.utils import distributed as dutil

    topo = dutil.setup_dist_util(DictConfig(dict(
        data_parallel_size=dp, tensor_parallel_size=tp, pipeline_parallel_size=pp,
        pipeline_num_layers=TINY["hidden_layers"], sequence_parallel=sp, device_type="cpu",
    )))
    assert (topo.data_parallel_size, topo.tensor_parallel_size, topo.pipeline_parallel_size) == (dp, tp, pp)
    with param_defaults(dtype=torch.float32, device="cpu", seed=7):
        model = GPTForPreTraining(DictConfig(TINY))
    opt = AdamW(get_default_optimizer_params(model, weight_decay_norm=0.0, weight_decay_bias=0.0,
                                             clip_grad_max_norm=1.0, clip_grad_norm_type=2.0), lr=1e-2, weight_decay=0.01)
    opt.configure(zero_stage=zero, param_names={id(p): n for n, p in model.named_parameters()})
    opt.setup()
    pipe = PipelineSchedule1F1B(model) if pp > 1 else None
    per_rank = GLOBAL_BATCH // dp
    micro = per_rank // acc
    losses = []
    for step in range(N_STEPS):
        ids, labels = _data(step)
        lo = topo.dp_rank * per_rank
        batches = [dict(input_ids=ids[lo + k * micro: lo + (k + 1) * micro], labels=labels[lo + k * micro: lo + (k + 1) * micro])
                   for k in range(acc)]
        opt.zero_grad()
        if pipe is not None:
            out = pipe.run(batches)
            loss = out["lm_loss"] if out is not None else torch.zeros(())
        else:
            loss = torch.zeros(())
            for b in batches:
                o = model(**b)
                (o["lm_loss"] / acc).backward()
                loss = loss + o["lm_loss"].detach() / acc
        opt.step()
        # mean over dp of the per-replica losses == global-batch mean loss
        v = loss.detach().clone().float()
        if pp > 1:
            import torch.distributed as dist
            dist.all_reduce(v, group=topo.pp_group)  # only the last stage holds a non-zero value
        v = dutil.dp_mean_to_rank0(v)
        losses.append(float(v))
    full = pstate.full_state_dict(model)
    if rank == 0:
        return losses, {k: v.clone() for k, v in full.items() if "tied_weight_copy" not in k}
    return None


def _single():
    return run_distributed(_train, 1, 1, 1, 1, False, 0, 1)[0]


@pytest.fixture(scope="module")
def baseline():
    return _single()


def _compare(base, other, tol=2e-5):
    bl, bp = base
    ol, op = other
    assert bl == pytest.approx(ol, rel=1e-4, abs=1e-5), (bl, ol)
    assert set(bp) == set(op)
    for k in bp:
        a, b = torch.as_tensor(bp[k]), torch.as_tensor(op[k])
        assert torch.allclose(a, b, rtol=1e-3, atol=tol), f"{k}: max diff {(a - b).abs().max()}"


@pytest.mark.parametrize(
    "world,dp,tp,pp,sp,zero,acc",
    [
        (2, 2, 1, 1, False, 0, 1),   # data parallel
        (2, 1, 2, 1, False, 0, 1),   # tensor parallel (all-reduce form, the reference's)
        (2, 1, 2, 1, True, 0, 1),    # tensor + sequence parallel (AG->GEMM / GEMM->RS form)
        (2, 1, 1, 2, False, 0, 4),   # pipeline 1F1B, 4 micro-batches, tied embeddings across stages
        (2, 2, 1, 1, False, 1, 2),   # ZeRO-1 + gradient accumulation
        (4, 2, 2, 1, True, 2, 1),    # dp x tp + SP + ZeRO-2
        (4, 1, 2, 2, False, 0, 2),   # tp x pp
        (4, 2, 1, 2, False, 1, 2),   # dp x pp + ZeRO
    ],
)
def test_layout_matches_single_process(baseline, world, dp, tp, pp, sp, zero, acc):
    res = run_distributed(_train, world, dp, tp, pp, sp, zero, acc)[0]
    _compare(baseline, res)


def test_gradient_accumulation_matches(baseline):
    res = run_distributed(_train, 1, 1, 1, 1, False, 0, 4)[0]
    _compare(baseline, res)


# -------------------------------------------------------------------------------------------------
def _linear_parity(rank, world):
    from libai_b200.config import DictConfig
    from libai_b200.layers import Linear
    from libai_b200.layers._param import param_defaults
    from libai_b200.parallel import state as pstate
    from libai_b200.utils import distributed as dutil

    dutil.setup_dist_util(DictConfig(dict(data_parallel_size=1, tensor_parallel_size=world, pipeline_parallel_size=1, device_type="cpu")))
    torch.manual_seed(0)
    x = torch.randn(5, 8, requires_grad=True)
    out = {}
    for mode in ("data", "col", "row"):
        with param_defaults(dtype=torch.float32, device="cpu", seed=3):
            lin = Linear(8, 12, parallel=mode)
        w = pstate.gather_tp(lin.weight.detach(), lin.weight.tp_dim)
        b = pstate.gather_tp(lin.bias.detach(), lin.bias.tp_dim)
        ref = torch.nn.functional.linear(x, w, b)
        if mode == "row":
            y = lin(pstate.shard_tp(x, 1))
        else:
            y = lin(x)
        if mode == "col":
            from libai_b200.parallel.mappings import gather_from_tp

            y = gather_from_tp(y)
        out[mode] = float((y - ref).abs().max())
        y.sum().backward()
    return out


def test_linear_col_row_match_dense():
    for res in run_distributed(_linear_parity, 2):
        for mode, err in res.items():
            assert err < 1e-5, (mode, err)


def test_layer_stage_map_goldens():
    from libai_b200.utils.distributed import compute_layer_stage_ids

    assert compute_layer_stage_ids(24, 4) == [0] * 6 + [1] * 7 + [2] * 7 + [3] * 4
    assert compute_layer_stage_ids(6, 2) == [0, 0, 0, 1, 1, 1]
    assert compute_layer_stage_ids(7, 2) == [0, 0, 0, 1, 1, 1, 1]
    assert compute_layer_stage_ids(12, 4) == [0] * 3 + [1] * 4 + [2] * 4 + [3] * 1
    assert compute_layer_stage_ids(5, 1) == [0] * 5


def _coords(rank, world):
    from libai_b200.config import DictConfig
    from libai_b200.utils import distributed as dutil

    t = dutil.setup_dist_util(DictConfig(dict(data_parallel_size=2, tensor_parallel_size=2, pipeline_parallel_size=2,
                                              pipeline_num_layers=4, device_type="cpu")))
    return (t.pp_rank, t.dp_rank, t.tp_rank, t.tp_ranks, t.dp_ranks, t.pp_ranks)


def test_rank_mesh_pp_outer_dp_middle_tp_inner():
    res = run_distributed(_coords, 8)
    for rank, (pp, dp, tp, tpr, dpr, ppr) in enumerate(res):
        assert rank == pp * 4 + dp * 2 + tp
        assert tpr == [pp * 4 + dp * 2 + i for i in range(2)]
        assert dpr == [pp * 4 + j * 2 + tp for j in range(2)]
        assert ppr == [s * 4 + dp * 2 + tp for s in range(2)]
