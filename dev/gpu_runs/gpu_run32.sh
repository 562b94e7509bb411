set -x
mkdir -p gpurun_out
timeout 900 python tests/gpu_kernel_check.py --out gpurun_out/kc32.json > gpurun_out/kc32.log 2>&1; echo "kc rc=$?"
grep -E "SUMMARY|\"ok\": false|bias\+residual|Traceback" gpurun_out/kc32.log | cut -c1-900
timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -n 3 | cut -c1-300
timeout 600 python bench.py > gpurun_out/bench_r32.log 2>&1; echo "bench rc=$?"; tail -n 1 gpurun_out/bench_r32.log | cut -c1-1600
