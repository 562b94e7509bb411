"""Tensor-parallel inference of the HuggingFace bloom implementation (reference
projects/mock_transformers/dist_infer_bloom.py): the HF modelling code is used as is, its projections are swapped
for this framework's column/row-parallel layers by ``init_env.parallelize``.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 \
        projects/mock_transformers/dist_infer_bloom.py --model bigscience/bloom-560m
"""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))

from projects.mock_transformers._common import run  # noqa: E402

if __name__ == "__main__":
    run("bloom", "bigscience/bloom-560m", "Hello, I'm a language model,")
