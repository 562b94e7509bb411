set -x
mkdir -p gpurun_out
timeout 60 python tests/gpu_kernel_check.py --only "fused bias grad,mlp fused" --out gpurun_out/r38_fused_bias_grad.json 2>&1 | tail -n 4 | cut -c1-1500
timeout 60 python bench.py --steps 10 --warmup 3 --no-e2e > gpurun_out/r38_bench_default.log 2>&1; echo "bench rc=$?"; tail -n 1 gpurun_out/r38_bench_default.log | cut -c1-1200
timeout 60 python bench.py --steps 10 --warmup 3 --no-e2e --fused-bias-grad 1 > gpurun_out/r38_bench_fbg.log 2>&1; echo "bench fbg rc=$?"; tail -n 1 gpurun_out/r38_bench_fbg.log | cut -c1-1200
timeout 150 python -m pytest tests/ -x -q -m gpu -k "not bench_contract and not smoke_entry" 2>&1 | tail -n 4 | cut -c1-400
