set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/ -q -m gpu 2>&1 | tail -25
