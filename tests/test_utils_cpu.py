"""Small utilities: history buffer, timers, event storage + JSON writer, async file IO, download cache helpers."""
import json
import os
import time

import pytest

from libai_b200.utils.events import EventStorage, JSONWriter, get_event_storage
from libai_b200.utils.file_utils import cached_path, filename_to_url, get_md5, url_to_filename
from libai_b200.utils.history_buffer import HistoryBuffer
from libai_b200.utils.non_blocking_io import NonBlockingIOManager
from libai_b200.utils.timer import Timer


def test_history_buffer_statistics():
    h = HistoryBuffer(max_length=4)
    for i, v in enumerate([1.0, 2.0, 3.0, 10.0, 4.0]):
        h.update(v, iteration=i)
    assert h.latest() == 4.0
    assert [v for v, _ in h.values()] == [2.0, 3.0, 10.0, 4.0]          # capped window
    assert h.median(3) == 4.0 and h.avg(2) == 7.0
    assert h.global_avg() == pytest.approx(4.0)                          # over everything ever seen


def test_timer_pause_resume():
    t = Timer()
    time.sleep(0.02)
    t.pause()
    frozen = t.seconds()
    assert t.is_paused() and frozen >= 0.015
    time.sleep(0.02)
    assert t.seconds() == pytest.approx(frozen)
    t.resume()
    time.sleep(0.01)
    assert t.seconds() > frozen
    t.reset()
    assert t.seconds() < frozen


def test_event_storage_and_json_writer(tmp_path):
    path = tmp_path / "m" / "metrics.json"
    writer = JSONWriter(str(path), window_size=2)
    with EventStorage(start_iter=5) as storage:
        assert get_event_storage() is storage
        for k in range(3):
            storage.put_scalar("loss", 4.0 - k)
            storage.put_scalars(lr=0.1, smoothing_hint=False)
            writer.write()
            storage.step()
        assert storage.iter == 8 and storage.history("loss").latest() == 2.0
        with storage.name_scope("eval"):
            storage.put_scalar("acc", 0.5)
        assert "eval/acc" in storage.latest()
    writer.close()
    rows = [json.loads(ln) for ln in open(path)]
    assert [r["iteration"] for r in rows] == [5, 6, 7]
    assert rows[0]["loss"] == 4.0 and rows[2]["loss"] == pytest.approx(2.5)   # median-smoothed over the window
    assert all(r["lr"] == 0.1 for r in rows)
    with pytest.raises(AssertionError):
        get_event_storage()


def test_non_blocking_io_orders_writes_and_joins(tmp_path):
    mgr = NonBlockingIOManager(buffered=False)
    path = str(tmp_path / "out.txt")
    closed = []
    f = mgr.get_non_blocking_io(path, open(path, "w"), callback_after_file_close=lambda: closed.append(True))
    for i in range(200):
        f.write(f"{i}\n")
    f.close()
    assert mgr._join(path) and closed == [True]
    assert open(path).read().split() == [str(i) for i in range(200)]
    with pytest.raises(ValueError):
        mgr._join(str(tmp_path / "never_opened"))
    assert mgr._close_thread_pool()


def test_cache_helpers(tmp_path):
    name = url_to_filename("https://example.org/a.bin", etag="v1")
    assert len(name) == 64 + 1 + 64 and name != url_to_filename("https://example.org/a.bin")
    local = tmp_path / name
    local.write_bytes(b"hello")
    with open(str(local) + ".json", "w") as fh:
        json.dump({"url": "https://example.org/a.bin", "etag": "v1"}, fh)
    assert filename_to_url(name, cache_dir=tmp_path) == ("https://example.org/a.bin", "v1")
    assert cached_path(str(local)) == str(local)
    assert get_md5(str(local)) == "5d41402abc4b2a76b9719d911017c592"
    with pytest.raises(FileNotFoundError):
        cached_path(str(tmp_path / "missing"))
    with pytest.raises(ValueError):
        cached_path("ftp://example.org/x")


def test_placement_and_sbp_vocabulary():
    """reference libai/utils/distributed.py:317-393 (get_layer_placement / get_nd_sbp / get_hidden_sbp / same_sbp)."""
    from libai_b200.config import DictConfig
    from libai_b200.utils import distributed as dist

    dist.reset_dist_util()
    try:
        dist.setup_dist_util(DictConfig(dict(data_parallel_size=1, tensor_parallel_size=1, pipeline_parallel_size=1,
                                             pipeline_num_layers=4, device_type="cpu")))
        p = dist.get_layer_placement(0)
        assert p.ranks == (0,) and p.mesh == ((0,),) and p.is_local and 0 in p and p.device.type == "cpu"
        assert dist.get_layer_placement(-1) == p and "ranks=[0]" in repr(p)
        assert dist.get_nd_sbp(["split_0", "broadcast"]) == ["broadcast"]      # single device: everything replicated
        assert dist.get_hidden_sbp() == ["broadcast"]
        assert dist.same_sbp(["split_0", "broadcast"], ["split_0", "broadcast"])
        assert not dist.same_sbp(["split_0", "broadcast"], ["broadcast", "broadcast"])
        # layouts (no process groups needed to ask the questions): fake a dp2 x tp2 x pp2 topology object
        topo = dist.get_dist_util()
        topo.data_parallel_size, topo.tensor_parallel_size, topo.pipeline_parallel_size = 2, 2, 2
        topo._layer_stage_ids = [0, 0, 1, 1]
        assert dist.get_nd_sbp(["split_0", "split_1"]) == ["split_0", "split_1"]
        assert dist.get_layer_placement(3).mesh == ((4, 5), (6, 7)) and dist.get_layer_placement(0).ranks == (0, 1, 2, 3)
        assert not dist.get_layer_placement(3).is_local
        topo.tensor_parallel_size = 1
        assert dist.get_nd_sbp(["split_0", "split_1"]) == ["split_0"]
        topo.data_parallel_size, topo.tensor_parallel_size = 1, 2
        assert dist.get_nd_sbp(["split_0", "split_1"]) == ["split_1"]
    finally:
        dist.reset_dist_util()


def test_reference_named_helpers_exist_and_work(tmp_path):
    """Names a reference user imports directly: instantiate_cfg, ModelLoader, s3_request, gpt2Graph / t5Graph."""
    import torch

    from libai_b200.config import LazyCall
    from libai_b200.config.instantiate import instantiate_cfg
    from libai_b200.models.utils.model_loader.base_loader import LoadPretrainedBase, ModelLoader
    from libai_b200.onnx_export.gpt2_to_onnx import gpt2Graph
    from libai_b200.onnx_export.t5_to_onnx import t5Graph
    from libai_b200.utils.file_utils import s3_request

    lin = instantiate_cfg(LazyCall(torch.nn.Linear)(in_features=3, out_features=2))
    assert isinstance(lin, torch.nn.Linear) and lin.out_features == 2
    node = LazyCall(dict)(a=LazyCall(torch.nn.ReLU)())
    assert isinstance(instantiate_cfg(node)["a"], torch.nn.ReLU)
    assert not isinstance(instantiate_cfg(node, recursive=False)["a"], torch.nn.ReLU)     # children left as configs
    assert ModelLoader is LoadPretrainedBase

    class NotFound(Exception):
        response = {"Error": {"Code": "404"}}

    @s3_request
    def fetch(url):
        raise NotFound()

    try:
        fetch("s3://bucket/key")
        raise AssertionError("expected EnvironmentError")
    except EnvironmentError as e:
        assert "s3://bucket/key" in str(e)

    class Toy(torch.nn.Module):
        def forward(self, input_ids):
            return {"prediction_scores": input_ids.float() * 2}

    assert torch.equal(gpt2Graph(Toy())(torch.ones(1, 3, dtype=torch.long)), torch.full((1, 3), 2.0))
    assert t5Graph(Toy()).input_names[0] == "encoder_input_ids"
