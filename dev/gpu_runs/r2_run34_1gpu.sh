# round 2, run 34 (1 GPU): LayerNorm backward adds its column partials straight into main_grad (no colreduce launch)
set -x
mkdir -p gpurun_out
timeout 600 python tests/gpu_kernel_check.py --quick --out gpurun_out/r2_34_kernel_check.json > gpurun_out/r2_34_kernel_check.log 2>&1; echo "kernel check rc=$?"; tail -1 gpurun_out/r2_34_kernel_check.log | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 --ref-same-box 0 --no-e2e > gpurun_out/r2_34_bench_1gpu.json 2> gpurun_out/r2_34_bench.err; tail -1 gpurun_out/r2_34_bench_1gpu.json | cut -c1-300
timeout 900 python -m pytest tests/ -x -q -m gpu > gpurun_out/r2_34_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_34_pytest_gpu.log
