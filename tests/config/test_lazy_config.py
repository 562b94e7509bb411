"""LazyConfig load / save / overrides / to_py (test model: reference tests/config/test_lazy_config.py)."""
import os
import tempfile
from itertools import count

from libai_b200.config import LazyCall, LazyConfig
from libai_b200.config import DictConfig


def root_filename():
    return os.path.join(os.path.dirname(__file__), "root_cfg.py")


def test_load():
    cfg = LazyConfig.load(root_filename())
    assert cfg.dir1a_dict.a == "modified"
    assert cfg.dir1b_dict.a == 1
    assert cfg.lazyobj.x == "base_a_1"
    cfg.lazyobj.x = "new_x"
    # reload
    cfg = LazyConfig.load(root_filename())
    assert cfg.lazyobj.x == "base_a_1"


def test_save_load():
    cfg = LazyConfig.load(root_filename())
    with tempfile.TemporaryDirectory(prefix="libai_b200") as d:
        fname = os.path.join(d, "test_config.yaml")
        LazyConfig.save(cfg, fname)
        cfg2 = LazyConfig.load(fname)
    assert cfg2.lazyobj._target_ == "itertools.count"
    assert cfg.lazyobj._target_ == count
    cfg2.lazyobj.pop("_target_")
    cfg.lazyobj.pop("_target_")
    # the rest are equal
    assert cfg == cfg2


def test_overrides():
    cfg = LazyConfig.load(root_filename())
    LazyConfig.apply_overrides(cfg, ["lazyobj.x=123", 'dir1b_dict.a="123"'])
    assert cfg.dir1b_dict.a == "123"
    assert cfg.lazyobj.x == 123
    LazyConfig.apply_overrides(cfg, ["dir1b_dict.new_key=[1,2]", "lazyobj.y=null", "dir1b_dict.flag=true"])
    assert list(cfg.dir1b_dict.new_key) == [1, 2] and cfg.lazyobj.y is None and cfg.dir1b_dict.flag is True


def test_invalid_overrides():
    cfg = LazyConfig.load(root_filename())
    try:
        LazyConfig.apply_overrides(cfg, ["lazyobj.x.xxx=123"])
    except Exception:
        return
    raise AssertionError("overriding below a leaf must fail")


def test_to_py():
    cfg = LazyConfig.load(root_filename())
    cfg.lazyobj.x = {"a": 1, "b": 2, "c": LazyCall(count)(x={"r": "a", "s": 2.4, "t": [1, 2, 3, "z"]})}
    cfg.list = ["a", 1, "b", 3.2]
    py_str = LazyConfig.to_py(cfg)
    assert "cfg.lazyobj = itertools.count(" in py_str and "cfg.list = ['a', 1, 'b', 3.2]" in py_str.replace('"', "'")
    assert "cfg.dir1a_dict.a = " in py_str


def test_packaged_configs():
    from libai_b200.config import get_config

    train = get_config("common/train.py").train
    assert "dist" in train and train.dist.pipeline_parallel_size == 1
    optim = get_config("common/optim.py").optim
    assert optim.lr > 0
