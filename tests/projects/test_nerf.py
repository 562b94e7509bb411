"""NeRF project: sampling maths, optimizers, datasets, and the config through the trainer on the analytic scene."""
import json
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(REPO, "tools"))


def test_sample_pdf_follows_weights():
    from projects.NeRF.modeling.system import sample_pdf

    bins = torch.linspace(0, 1, 11)[None].repeat(2, 1)
    w = torch.zeros(2, 10)
    w[0, 3] = 1.0
    w[1, 7] = 1.0
    s = sample_pdf(bins, w, 64, det=True)
    assert ((s[0] > 0.29) & (s[0] < 0.41)).float().mean() > 0.95
    assert ((s[1] > 0.69) & (s[1] < 0.81)).float().mean() > 0.95
    torch.manual_seed(0)
    u = sample_pdf(bins, torch.ones(2, 10), 4096, det=False)
    assert abs(u.mean().item() - 0.5) < 0.02 and u.min() >= 0 and u.max() <= 1


def test_embedding_layout():
    from projects.NeRF.modeling.nerf import Embedding

    e = Embedding(3, 4)
    x = torch.randn(5, 3)
    y = e(x)
    assert y.shape == (5, 27) and torch.equal(y[:, :3], x)
    assert torch.allclose(y[:, 3:6], torch.sin(x)) and torch.allclose(y[:, 6:9], torch.cos(x))
    assert torch.allclose(y[:, 9:12], torch.sin(2 * x))


def test_radam_matches_torch_and_ranger_runs():
    from projects.NeRF.optimizers import RAdam, Ranger

    torch.manual_seed(0)
    w0 = torch.randn(7, 5)
    a, b = torch.nn.Parameter(w0.clone()), torch.nn.Parameter(w0.clone())
    oa, ob = RAdam([a], lr=1e-2), torch.optim.RAdam([b], lr=1e-2)
    for i in range(12):
        g = torch.randn(7, 5, generator=torch.Generator().manual_seed(i))
        a.grad, b.grad = g.clone(), g.clone()
        oa.step()
        ob.step()
    assert (a - b).abs().max() < 1e-5
    c = torch.nn.Parameter(w0.clone())
    oc = Ranger([c], lr=1e-2, k=3, use_gc=True)
    for i in range(9):
        c.grad = 2 * c.detach()
        oc.step()
    assert c.abs().sum() < w0.abs().sum() and "slow_buffer" in oc.state[c]


def test_analytic_dataset_geometry():
    from projects.NeRF.datasets.nerf_dataset import AnalyticSceneDataset, get_ndc_rays

    ds = AnalyticSceneDataset(split="val", img_wh=(16, 16))
    it = ds[0]
    rays, rgbs, mask = it.get("rays").tensor, it.get("rgbs").tensor, it.get("valid_mask").tensor
    assert rays.shape == (256, 8) and rgbs.shape == (256, 3)
    assert torch.allclose(rays[:, :3].norm(dim=-1), torch.full((256,), 4.0), atol=1e-4)     # cameras on r=4 sphere
    centre = rays[16 * 8 + 8]
    d = centre[3:6] / centre[3:6].norm()
    assert torch.allclose(-d, centre[:3] / 4.0, atol=0.1)                                   # looks at the origin
    assert mask.view(16, 16)[8, 8] and not mask.view(16, 16)[0, 0]
    assert (rgbs[~mask] == 1).all()
    tr = AnalyticSceneDataset(split="train", img_wh=(16, 16), batchsize=32)
    first = tr[0]
    assert first.get("rays").tensor.shape == (32, 8) and len(tr) == 8 * 256 // 32
    o, dd = get_ndc_rays(4, 4, 2.0, 1.0, torch.tensor([[0.0, 0, 0]]), torch.tensor([[0.0, 0, -1]]))
    assert torch.allclose(o, torch.tensor([[0.0, 0, -1]])) and torch.allclose(dd, torch.tensor([[0.0, 0, 2]]))


def test_llff_dataset(tmp_path):
    from PIL import Image

    from projects.NeRF.datasets.nerf_dataset import LLFFDataset

    root = tmp_path / "scene"
    (root / "images").mkdir(parents=True)
    rng = np.random.default_rng(0)
    rows = []
    for i in range(4):
        Image.fromarray(rng.integers(0, 255, (12, 16, 3), dtype=np.uint8)).save(root / "images" / f"{i:03d}.png")
        pose = np.concatenate([np.eye(3), np.array([[0.1 * i], [0.0], [0.0]]), np.array([[12.0], [16.0], [20.0]])], 1)
        rows.append(np.concatenate([pose.reshape(-1), [2.0, 9.0]]))
    np.save(root / "poses_bounds.npy", np.stack(rows))
    tr = LLFFDataset(str(root), "train", img_wh=(16, 12), batchsize=8)
    assert tr.all_rays.shape == (3, 192, 8) and tr[0].get("rgbs").tensor.shape == (8, 3)
    assert float(tr.all_rays[..., 6].max()) == 0.0 and float(tr.all_rays[..., 7].min()) == 1.0   # NDC near/far
    va = LLFFDataset(str(root), "val", img_wh=(16, 12))
    assert len(va) == 1 and va[0].get("rays").tensor.shape == (192, 8)
    te = LLFFDataset(str(root), "test", img_wh=(16, 12))
    assert len(te) == 120 and "rgbs" not in te[0].get_fields()


def test_train_eval_through_trainer(tmp_path, monkeypatch):
    monkeypatch.setenv("NERF_DATASET_TYPE", "Analytic")
    import train_net
    from libai_b200.config import default_argument_parser

    out = str(tmp_path / "nerf")
    argv = ["--config-file", os.path.join(REPO, "projects/NeRF/configs/config_nerf.py"), "model.cfg.W=32",
            "model.cfg.N_samples=8", "model.cfg.N_importance=8", "train.rays_per_batch=128", "train.train_iter=30",
            "train.train_epoch=0", "train.amp.enabled=false", "train.log_period=1", "train.evaluation.eval_period=30",
            "dataloader.train.dataset.0.img_wh=[12,12]", "dataloader.train.dataset.0.batchsize=128",
            "dataloader.test.0.dataset.img_wh=[12,12]", "dataloader.test.0.dataset.n_views=2",
            "train.evaluation.evaluator.img_wh=[12,12]", "optim.lr=5e-3", f"train.output_dir={out}"]
    train_net.main(default_argument_parser().parse_args(argv))
    metrics = [json.loads(ln) for ln in open(os.path.join(out, "metrics.json"))]
    losses = [m["total_loss"] for m in metrics if "total_loss" in m]
    assert len(losses) >= 30 and losses[-1] < losses[0]
    assert any("psnr" in m for m in metrics)
