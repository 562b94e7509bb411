"""PathManager / LazyPath / file_lock (model: reference tests/test_file_io.py)."""
import os

from libai_b200.utils.file_io import LazyPath, PathManager, file_lock


def test_native_handler_roundtrip(tmp_path):
    d = str(tmp_path / "a" / "b")
    PathManager.mkdirs(d)
    assert PathManager.isdir(d) and PathManager.exists(d)
    fp = os.path.join(d, "x.txt")
    with PathManager.open(fp, "w") as f:
        f.write("hello")
    assert PathManager.isfile(fp) and PathManager.get_local_path(fp) == fp
    with PathManager.open(fp, "r") as f:
        assert f.read() == "hello"
    PathManager.copy(fp, fp + ".copy", overwrite=True)
    assert sorted(PathManager.ls(d)) == ["x.txt", "x.txt.copy"]
    PathManager.rm(fp + ".copy")
    assert not PathManager.exists(fp + ".copy")


def test_lazy_path_resolves_once():
    calls = []

    def resolve():
        calls.append(1)
        return "/tmp/some/where"

    p = LazyPath(resolve)
    assert not calls
    assert os.fspath(p) == "/tmp/some/where" and str(p).endswith("where") and p.startswith("/tmp")
    assert len(calls) == 1


def test_file_lock(tmp_path):
    target = str(tmp_path / "artifact")
    with file_lock(target):
        open(target, "w").write("x")
    assert os.path.exists(target)
