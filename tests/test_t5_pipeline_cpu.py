"""T5 under pipeline parallelism: the (encoder stream, decoder stream) pair travels between stages; loss and updated
weights after three 1F1B steps equal the single-process run (tied embedding across first/last stage included)."""
import pytest
import torch

from tests.dist_utils import run_distributed

T5_TINY = dict(
    vocab_size=96, hidden_size=32, hidden_layers=2, num_attention_heads=4, intermediate_size=64,
    embedding_dropout_prob=0.0, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, max_position_embeddings=16,
    initializer_range=0.02, layernorm_eps=1e-5, bias_gelu_fusion=True, bias_dropout_fusion=True,
    scale_mask_softmax_fusion=True, apply_query_key_layer_scaling=True, apply_residual_post_layernorm=False,
    amp_enabled=False,
)
B, SE, SD, STEPS = 8, 12, 8, 3


def _batch(step):
    g = torch.Generator().manual_seed(50 + step)
    enc = torch.randint(1, 96, (B, SE), generator=g)
    dec = torch.randint(1, 96, (B, SD), generator=g)
    lab = torch.randint(1, 96, (B, SD), generator=g)
    enc_len = torch.randint(SE // 2, SE + 1, (B,), generator=g)
    em = (torch.arange(SE)[None] < enc_len[:, None])
    return dict(
        encoder_input_ids=enc, decoder_input_ids=dec,
        encoder_attn_mask=(em[:, :, None] & em[:, None, :]),
        decoder_attn_mask=torch.ones(B, SD, SD, dtype=torch.bool).tril(),
        encoder_decoder_attn_mask=em[:, None, :].expand(B, SD, SE).contiguous(),
        lm_labels=lab, loss_mask=torch.ones(B, SD),
    )


def _train(rank, world, pp, acc):
    from libai_b200.config import DictConfig
    from libai_b200.layers._param import param_defaults
    from libai_b200.models.t5_model import T5ForPreTraining
    from libai_b200.optim import AdamW, get_default_optimizer_params
    from libai_b200.parallel import state as pstate
    from libai_b200.parallel.pipeline import PipelineSchedule1F1B
    from libai_b200.utils import distributed as dutil
    import torch.distributed as dist

    topo = dutil.setup_dist_util(DictConfig(dict(
        data_parallel_size=1, tensor_parallel_size=1, pipeline_parallel_size=pp,
        pipeline_num_layers=2 * T5_TINY["hidden_layers"], device_type="cpu")))
    with param_defaults(dtype=torch.float32, device="cpu", seed=3):
        model = T5ForPreTraining(DictConfig(T5_TINY))
    opt = AdamW(get_default_optimizer_params(model, clip_grad_max_norm=1.0, clip_grad_norm_type=2.0), lr=1e-2, weight_decay=0.01)
    opt.configure(zero_stage=0, param_names={id(p): n for n, p in model.named_parameters()})
    opt.setup()
    pipe = PipelineSchedule1F1B(model) if pp > 1 else None
    micro = B // acc
    losses = []
    for step in range(STEPS):
        full = _batch(step)
        batches = [{k: v[i * micro:(i + 1) * micro] for k, v in full.items()} for i in range(acc)]
        opt.zero_grad()
        if pipe is not None:
            out = pipe.run(batches)
            loss = out["masked_lm_loss"] if out is not None else torch.zeros(())
        else:
            loss = torch.zeros(())
            for b in batches:
                o = model.forward_stage(b)
                (o["masked_lm_loss"] / acc).backward()
                loss = loss + o["masked_lm_loss"].detach() / acc
        opt.step()
        v = loss.detach().clone().float()
        if pp > 1:
            dist.all_reduce(v, group=topo.pp_group)
        losses.append(float(v))
    state = pstate.full_state_dict(model)
    if rank == 0:
        return losses, {k: v.clone() for k, v in state.items() if "tied_weight_copy" not in k}
    return None


def _staged_vs_plain(rank, world):
    from libai_b200.config import DictConfig
    from libai_b200.layers._param import param_defaults
    from libai_b200.models.t5_model import T5ForPreTraining
    from libai_b200.utils import distributed as dutil

    dutil.setup_dist_util(DictConfig(dict(data_parallel_size=1, tensor_parallel_size=1, pipeline_parallel_size=1,
                                          device_type="cpu")))
    with param_defaults(dtype=torch.float32, device="cpu", seed=3):
        model = T5ForPreTraining(DictConfig(T5_TINY))
    b = _batch(0)
    return float(model(**b)["masked_lm_loss"]), float(model.forward_stage(b)["masked_lm_loss"])


def test_staged_forward_equals_plain_forward():
    a, b = run_distributed(_staged_vs_plain, 1)[0]
    assert a == pytest.approx(b, rel=1e-6)


def test_pipeline_matches_single_process():
    base_l, base_p = run_distributed(_train, 1, 1, 2)[0]
    pp_l, pp_p = run_distributed(_train, 2, 2, 4)[0]
    assert base_l == pytest.approx(pp_l, rel=1e-4, abs=1e-5), (base_l, pp_l)
    assert set(base_p) == set(pp_p)
    for k in base_p:
        a, b = torch.as_tensor(base_p[k]), torch.as_tensor(pp_p[k])
        assert torch.allclose(a, b, rtol=1e-3, atol=2e-5), f"{k}: {(a - b).abs().max()}"
