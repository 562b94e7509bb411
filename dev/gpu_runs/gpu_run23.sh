set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu.py -q -x -k "cuda_graph" 2>&1 | tail -n 15
timeout 600 python bench.py --steps 10 --warmup 3 --graphs 1 > gpurun_out/bench_r23_graphs.log 2>&1; echo "bench graphs rc=$?"; tail -n 3 gpurun_out/bench_r23_graphs.log | cut -c1-1500
timeout 600 python tools/host_profile.py --steps 6 > gpurun_out/host_profile23.log 2>&1; echo "host profile rc=$?"; grep -A48 "host enqueue" gpurun_out/host_profile23.log | cut -c1-160 | head -64
