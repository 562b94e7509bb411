set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu.py -q -x -k "cuda_graph" 2>&1 | tail -n 12 | cut -c1-400
timeout 600 python bench.py --steps 10 --warmup 3 --graphs 1 > gpurun_out/bench_r24_graphs.log 2>&1; echo "bench graphs rc=$?"; grep -n "Error" gpurun_out/bench_r24_graphs.log | head -5; tail -n 1 gpurun_out/bench_r24_graphs.log | cut -c1-1600
