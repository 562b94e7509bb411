set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv
timeout 420 python tests/gpu_kernel_check.py --only gemm --out gpurun_out/kc_gemm.json > gpurun_out/kc_gemm.log 2>&1; echo "gemm rc=$?"
timeout 300 python tests/gpu_kernel_check.py --only linear,norm,layernorm,elementwise,rope,cross,fused --out gpurun_out/kc_misc.json > gpurun_out/kc_misc.log 2>&1; echo "misc rc=$?"
timeout 300 python tests/gpu_kernel_check.py --only attention --out gpurun_out/kc_attn.json > gpurun_out/kc_attn.log 2>&1; echo "attn rc=$?"
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
LIBAI_B200_IMPL=ref timeout 600 python bench.py --impl ref --steps 5 --warmup 3 --no-e2e > gpurun_out/bench_ref.log 2>&1; echo "bench_ref rc=$?"
tail -5 gpurun_out/kc_gemm.log gpurun_out/kc_misc.log gpurun_out/kc_attn.log gpurun_out/smoke.log gpurun_out/bench.log gpurun_out/bench_ref.log
