set -x
mkdir -p gpurun_out
timeout 600 python tests/gpu_kernel_check.py --only "accumulating,norm,layernorm" --out gpurun_out/kc27.json > gpurun_out/kc27.log 2>&1; echo "kc rc=$?"
grep -E "SUMMARY|\"ok\": false|accumulating" gpurun_out/kc27.log | cut -c1-700
timeout 900 python -m pytest tests/test_gpu.py -q -x -k "not bench_contract and not nvlink" 2>&1 | tail -n 5 | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r27.log 2>&1; echo "bench rc=$?"; tail -n 1 gpurun_out/bench_r27.log | cut -c1-1500
