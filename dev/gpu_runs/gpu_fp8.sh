set -x
mkdir -p gpurun_out
timeout 240 python tests/gpu_kernel_check.py --only "fp8" --out gpurun_out/r35_fp8_kernel_check.json 2>&1 | tail -n 30 | cut -c1-1500
timeout 400 python -m pytest tests/test_gpu.py -x -q -k "fp8 or cuda_graph or native_training_step" 2>&1 | tail -n 15 | cut -c1-600
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r35_bench_bf16.log 2>&1; echo "bench rc=$?"; tail -n 1 gpurun_out/r35_bench_bf16.log | cut -c1-1500
timeout 300 python bench.py --steps 10 --warmup 3 --fp8 1 --no-e2e > gpurun_out/r35_bench_fp8.log 2>&1; echo "bench fp8 rc=$?"; tail -n 1 gpurun_out/r35_bench_fp8.log | cut -c1-1500
