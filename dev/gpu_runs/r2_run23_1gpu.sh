# round 2, run 23 (1 GPU): bandwidth-oriented LayerNorm backward + wider forward grid: numerics, speed, step
set -x
mkdir -p gpurun_out
timeout 600 python tests/gpu_kernel_check.py --out gpurun_out/r2_23_kernel_check.json > gpurun_out/r2_23_kernel_check.log 2>&1; echo "kernel check rc=$?"; tail -1 gpurun_out/r2_23_kernel_check.log | cut -c1-400
grep -h "layernorm speed\|norm rms" gpurun_out/r2_23_kernel_check.log | cut -c1-300 | tail -8
timeout 600 python bench.py --steps 20 --warmup 5 --ref-same-box 0 --no-e2e > gpurun_out/r2_23_bench_1gpu.json 2> gpurun_out/r2_23_bench.err; tail -1 gpurun_out/r2_23_bench_1gpu.json | cut -c1-500
