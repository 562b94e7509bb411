# round 2, run 21 (1 GPU): the driver's GPU tier (pytest -m gpu) + full kernel check (bias / ALiBi / dropout variants of the
# pipelined attention backward) + smoke()
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/r2_21_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_21_pytest_gpu.log
timeout 600 python tests/gpu_kernel_check.py --out gpurun_out/r2_21_kernel_check.json > gpurun_out/r2_21_kernel_check.log 2>&1; echo "kernel check rc=$?"; tail -2 gpurun_out/r2_21_kernel_check.log | cut -c1-600
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_21_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2_21_smoke.log
