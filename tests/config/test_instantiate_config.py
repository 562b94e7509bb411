"""instantiate(): targets, nesting, lists, dataclasses, `_recursive_` (model: reference
tests/config/test_instantiate_config.py)."""
import os
import tempfile
from collections import namedtuple
from dataclasses import dataclass

import pytest

from libai_b200.config import DictConfig, LazyCall, LazyConfig, OmegaConf, instantiate


class ShapeSpec(namedtuple("_ShapeSpec", ["channels", "width"])):
    def __new__(cls, channels=None, width=None):
        return super().__new__(cls, channels, width)


class Holder:
    def __init__(self, int_arg, list_arg=None, dict_arg=None, extra_arg=None):
        self.int_arg, self.list_arg, self.dict_arg, self.extra_arg = int_arg, list_arg, dict_arg, extra_arg

    def __call__(self, call_arg):
        return call_arg + self.int_arg


@dataclass
class Spec:
    channels: int = 1
    width: int = 3


def test_basic_construct():
    objconf = LazyCall(Holder)(
        int_arg=3, list_arg=[10], dict_arg={}, extra_arg=LazyCall(Holder)(int_arg=4, list_arg="${..list_arg}")
    )
    obj = instantiate(objconf)
    assert isinstance(obj, Holder) and obj.int_arg == 3
    assert obj.extra_arg.int_arg == 4
    objconf.extra_arg.list_arg = [5]
    assert instantiate(objconf).extra_arg.list_arg == [5]


def test_instantiate_other_obj():
    assert instantiate(5) == 5
    x = [3, 4, 5]
    assert list(instantiate(x)) == x
    x = Holder(1)
    assert instantiate(x) is x
    assert instantiate({"xx": "yy"}) == {"xx": "yy"}


def test_instantiate_namedtuple_and_dataclass():
    x = LazyCall(Holder)(int_arg=ShapeSpec(channels=1, width=3))
    with tempfile.TemporaryDirectory() as d:
        LazyConfig.save(x, os.path.join(d, "x.yaml"))
    assert instantiate(x).int_arg.channels == 1
    y = LazyCall(Holder)(int_arg=LazyCall(Spec)(channels=7))
    assert instantiate(y).int_arg == Spec(7, 3)


def test_bad_lazycall():
    with pytest.raises(Exception):
        LazyCall(3)


def test_instantiate_lst_and_str_target():
    lst = [1, 2, LazyCall(Holder)(int_arg=1)]
    x = LazyCall(Holder)(int_arg=lst)
    y = instantiate(x)
    assert y.int_arg[0] == 1 and isinstance(y.int_arg[2], Holder)
    cfg = {"_target_": "collections.OrderedDict", "a": 1}
    import collections

    assert isinstance(instantiate(cfg), collections.OrderedDict)


def test_instantiate_no_recursive():
    def helper_func(obj):
        return isinstance(obj, (dict, DictConfig)) and "_target_" in obj

    objconf = LazyCall(helper_func)(obj=LazyCall(Holder)(int_arg=4))
    objconf["_recursive_"] = False
    assert instantiate(objconf) is True
    objconf["_recursive_"] = True
    assert instantiate(objconf) is False
