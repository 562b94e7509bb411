"""Every shipped config file parses with LazyConfig (library recipes and all projects)."""
import glob
import os

import pytest

from libai_b200.config import LazyConfig

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _configs():
    pats = ["configs/*.py", "configs/common/*.py", "configs/common/models/*.py", "configs/common/models/*/*.py",
            "configs/common/data/*.py", "projects/*/configs/*.py", "projects/*/configs/models/*.py", "projects/*/config/*.py",
            "projects/Llama/adapter/adapter_config.py", "projects/Llama/adapter/adapter_sft.py"]
    out = []
    for p in pats:
        out += [f for f in sorted(glob.glob(os.path.join(REPO, p))) if not f.endswith("__init__.py")]
    return [os.path.relpath(f, REPO) for f in out]


@pytest.mark.parametrize("path", _configs())
def test_config_loads(path, monkeypatch):
    monkeypatch.chdir(REPO)
    cfg = LazyConfig.load(path)
    assert len(cfg) > 0, path
