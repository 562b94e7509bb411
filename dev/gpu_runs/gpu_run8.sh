set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -15
timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_r8.log 2>&1; echo "bench rc=$?"; tail -n 1 gpurun_out/bench_r8.log | cut -c1-1500
