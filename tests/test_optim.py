"""Param-group expansion / reduction and optimizer parity (model: reference tests/test_optim.py)."""
import math

import torch

from libai_b200.optim import AdamW, SGD, get_default_optimizer_params
from libai_b200.optim.build import _expand_param_groups, reduce_param_groups


def test_expand_param_groups():
    params = [
        {"params": ["p1", "p2", "p3", "p4"], "lr": 1.0, "weight_decay": 3.0},
        {"params": ["p2", "p3", "p5"], "lr": 2.0, "momentum": 2.0},
        {"params": ["p1"], "weight_decay": 4.0},
    ]
    out = _expand_param_groups(params)
    gt = [
        dict(params=["p1"], lr=1.0, weight_decay=4.0),
        dict(params=["p2"], lr=2.0, weight_decay=3.0, momentum=2.0),
        dict(params=["p3"], lr=2.0, weight_decay=3.0, momentum=2.0),
        dict(params=["p4"], lr=1.0, weight_decay=3.0),
        dict(params=["p5"], lr=2.0, momentum=2.0),
    ]
    assert out == gt


def test_reduce_param_groups():
    params = [
        dict(params=["p1"], lr=1.0, weight_decay=4.0),
        dict(params=["p2", "p6"], lr=2.0, weight_decay=3.0, momentum=2.0),
        dict(params=["p3"], lr=2.0, weight_decay=3.0, momentum=2.0),
        dict(params=["p4"], lr=1.0, weight_decay=3.0),
        dict(params=["p5"], lr=2.0, momentum=2.0),
    ]
    gt = [
        {"lr": 1.0, "weight_decay": 4.0, "params": ["p1"]},
        {"lr": 2.0, "weight_decay": 3.0, "momentum": 2.0, "params": ["p2", "p6", "p3"]},
        {"lr": 1.0, "weight_decay": 3.0, "params": ["p4"]},
        {"lr": 2.0, "momentum": 2.0, "params": ["p5"]},
    ]
    assert reduce_param_groups(params) == gt


def test_default_params_norm_bias_overrides():
    model = torch.nn.Sequential(torch.nn.Linear(4, 4), torch.nn.LayerNorm(4))
    groups = get_default_optimizer_params(model, weight_decay=0.1, weight_decay_norm=0.0, weight_decay_bias=0.0,
                                          clip_grad_max_norm=1.0, clip_grad_norm_type=2.0)
    by_wd = {g["weight_decay"]: len(g["params"]) for g in groups}
    assert by_wd == {0.1: 1, 0.0: 3}
    assert all(g["clip_grad_max_norm"] == 1.0 for g in groups)


def _train(opt_cls, ref_cls, **kw):
    torch.manual_seed(0)
    a, b = torch.nn.Linear(6, 3), torch.nn.Linear(6, 3)
    b.load_state_dict(a.state_dict())
    oa = opt_cls([{"params": list(a.parameters())}], **kw)
    ob = ref_cls(b.parameters(), **kw)
    for step in range(4):
        x = torch.randn(5, 6)
        for m, o in ((a, oa), (b, ob)):
            o.zero_grad()
            m(x).pow(2).mean().backward()
            o.step()
    return max((p - q).abs().max().item() for p, q in zip(a.parameters(), b.parameters()))


def test_adamw_matches_torch():
    assert _train(AdamW, torch.optim.AdamW, lr=1e-2, weight_decay=0.01) < 1e-6


def test_sgd_matches_torch():
    assert _train(SGD, torch.optim.SGD, lr=1e-2, momentum=0.9) < 1e-6


def test_clip_and_state_dict_roundtrip():
    torch.manual_seed(0)
    m = torch.nn.Linear(4, 4)
    opt = AdamW([{"params": list(m.parameters()), "clip_grad_max_norm": 0.1, "clip_grad_norm_type": 2.0}], lr=1e-2)
    m(torch.randn(3, 4)).pow(2).sum().backward()
    opt.step()
    sd = opt.state_dict()
    m2 = torch.nn.Linear(4, 4)
    m2.load_state_dict(m.state_dict())
    opt2 = AdamW([{"params": list(m2.parameters()), "clip_grad_max_norm": 0.1, "clip_grad_norm_type": 2.0}], lr=1e-2)
    opt2.load_state_dict(sd)
    x = torch.randn(3, 4)
    for mm, oo in ((m, opt), (m2, opt2)):
        oo.zero_grad()
        mm(x).pow(2).sum().backward()
        oo.step()
    assert max((p - q).abs().max().item() for p, q in zip(m.parameters(), m2.parameters())) < 1e-7


def test_checkpointer_load_refreshes_fp32_master(tmp_path):
    """bf16 parameters + fp32 master: weights loaded through ``Checkpointer.load(..., checkpointables=[])`` (the
    ``train.load_weight`` path) must be what the first ``step`` updates, not the random-init master snapshot."""
    import torch
    from torch import nn

    from libai_b200.optim.optimizers import AdamW
    from libai_b200.utils.checkpoint import Checkpointer

    torch.manual_seed(0)
    model = nn.Linear(16, 16).to(torch.bfloat16)
    opt = AdamW(model.parameters(), lr=1e-3, weight_decay=0.0)
    opt.configure(zero_stage=0, param_names={id(p): n for n, p in model.named_parameters()})
    opt.setup()
    torch.save({"weight": torch.full((16, 16), 5.0), "bias": torch.full((16,), -3.0)}, tmp_path / "w.pt")
    Checkpointer(model, str(tmp_path), optimizer=opt).load(str(tmp_path / "w.pt"), checkpointables=[])
    assert float(model.weight.float().mean()) == 5.0
    opt.zero_grad()
    model(torch.randn(4, 16, dtype=torch.bfloat16)).float().sum().backward()
    opt.step()
    assert abs(float(model.weight.float().mean()) - 5.0) < 0.05 and abs(float(model.bias.float().mean()) + 3.0) < 0.05
