# round 2, run 28 (1 GPU): dq accumulator cleared by the delta kernel; attention numerics + timing, quick kernel check, bench
set -x
mkdir -p gpurun_out
timeout 200 python dev/attn_dev.py > gpurun_out/r2_28_attn.json 2> gpurun_out/r2_28_attn.err; echo "attn rc=$?"; tail -1 gpurun_out/r2_28_attn.json | cut -c1-1500
timeout 600 python tests/gpu_kernel_check.py --quick --out gpurun_out/r2_28_kernel_check.json > gpurun_out/r2_28_kernel_check.log 2>&1; echo "kernel check rc=$?"; tail -1 gpurun_out/r2_28_kernel_check.log | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 --ref-same-box 0 --no-e2e > gpurun_out/r2_28_bench_1gpu.json 2> gpurun_out/r2_28_bench.err; tail -1 gpurun_out/r2_28_bench_1gpu.json | cut -c1-300
LIBAI_B200_FUSED_BIAS_GRAD=1 timeout 600 python bench.py --steps 20 --warmup 5 --ref-same-box 0 --no-e2e > gpurun_out/r2_28_bench_1gpu_fused_bias_grad.json 2>> gpurun_out/r2_28_bench.err; tail -1 gpurun_out/r2_28_bench_1gpu_fused_bias_grad.json | cut -c1-300
