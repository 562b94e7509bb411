set -x
mkdir -p gpurun_out
timeout 120 python tools/gpu_debug_actgrad.py 2>&1 | tail -12
timeout 600 python tests/gpu_kernel_check.py --only "linear,mlp,accum,norm,elementwise,attention speed,gemm L1" --out gpurun_out/kc6.json > gpurun_out/kc6.log 2>&1; echo "kc rc=$?"
grep -E "SUMMARY|\"ok\": false" gpurun_out/kc6.log | cut -c1-400
grep -E "linear_fwd speed|mlp fused|attention speed|accumulating" gpurun_out/kc6.log | cut -c1-500
timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_r6.log 2>&1; echo "bench rc=$?"; tail -n 1 gpurun_out/bench_r6.log | cut -c1-1500
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_kernel -s 2 -c 1 -o gpurun_out/prof_attn_bwd python tests/gpu_kernel_check.py --only "attention speed" --out gpurun_out/tmp.json > gpurun_out/ncu_attn_bwd.log 2>&1; echo "ncu attn bwd rc=$?"
