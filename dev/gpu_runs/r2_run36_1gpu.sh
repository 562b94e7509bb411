# round 2, run 36 (1 GPU): final tree — driver GPU tier, smoke(), default bench
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -x -q -m gpu > gpurun_out/r2_36_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_36_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_36_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2_36_smoke.log
timeout 600 python bench.py > gpurun_out/r2_36_bench_1gpu_default.json 2> gpurun_out/r2_36_bench.err; echo "bench rc=$?"; tail -1 gpurun_out/r2_36_bench_1gpu_default.json | cut -c1-2500
