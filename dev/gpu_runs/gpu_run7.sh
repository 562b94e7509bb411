set -x
mkdir -p gpurun_out
timeout 120 python tools/gpu_debug_actgrad.py 2>&1 | tail -14
timeout 600 python tests/gpu_kernel_check.py --only "attention,accum,mlp,linear_fwd speed" --out gpurun_out/kc7.json > gpurun_out/kc7.log 2>&1; echo "kc rc=$?"
grep -E "SUMMARY|\"ok\": false" gpurun_out/kc7.log | cut -c1-400
grep -E "libai_b200\]|linear_fwd speed|attention speed|accumulating" gpurun_out/kc7.log | cut -c1-500
timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_r7.log 2>&1; echo "bench rc=$?"; tail -n 1 gpurun_out/bench_r7.log | cut -c1-1200
