"""Builds ``helpers.cpp`` into ``libai_b200/data/data_utils/_helpers.so`` (g++ -O3, C ABI)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "helpers.cpp")
SO = os.path.join(HERE, "_helpers.so")


def ensure_built(force: bool = False) -> str:
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(SRC):
        tmp = SO + f".tmp{os.getpid()}"
        subprocess.check_call(["g++", "-O3", "-std=c++17", "-shared", "-fPIC", "-o", tmp, SRC])
        os.replace(tmp, SO)
    return SO


if __name__ == "__main__":
    print(ensure_built(force=True))
