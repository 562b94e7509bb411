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
