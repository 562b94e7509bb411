# round 2, run 9 (1 GPU): cta_group::2 GEMM — numerics of every layout/epilogue, speed vs the single-CTA kernel and cuBLAS,
# then the training-step bench with both settings
set -x
mkdir -p gpurun_out
timeout 900 python tests/gpu_kernel_check.py --only "gemm,linear_fwd,mlp fused,fused bias grad,bias+residual GEMM" --out gpurun_out/r2_kernel_check_gemm_2cta.json > gpurun_out/r2_kernel_check_gemm_2cta.log 2>&1
tail -3 gpurun_out/r2_kernel_check_gemm_2cta.log | cut -c1-400
LIBAI_B200_GEMM_2CTA=0 timeout 900 python tests/gpu_kernel_check.py --only "gemm" --out gpurun_out/r2_kernel_check_gemm_1cta.json > gpurun_out/r2_kernel_check_gemm_1cta.log 2>&1
tail -2 gpurun_out/r2_kernel_check_gemm_1cta.log | cut -c1-400
timeout 600 python bench.py --gpus 1 --steps 15 --warmup 5 --ref-same-box 0 > gpurun_out/r2_bench_1gpu_2cta.json 2> gpurun_out/r2_bench_1gpu_2cta.err
cat gpurun_out/r2_bench_1gpu_2cta.json | cut -c1-700; tail -2 gpurun_out/r2_bench_1gpu_2cta.err | cut -c1-300
LIBAI_B200_GEMM_2CTA=0 timeout 600 python bench.py --gpus 1 --steps 15 --warmup 5 --ref-same-box 0 --no-e2e > gpurun_out/r2_bench_1gpu_1cta.json 2> gpurun_out/r2_bench_1gpu_1cta.err
cat gpurun_out/r2_bench_1gpu_1cta.json | cut -c1-500
