set -x
mkdir -p gpurun_out
timeout 600 python tests/gpu_kernel_check.py --only "attention" --out gpurun_out/kc10.json > gpurun_out/kc10.log 2>&1; echo "kc rc=$?"
grep -E "SUMMARY|\"ok\": false|attention speed" gpurun_out/kc10.log | cut -c1-500
timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_r10.log 2>&1; echo "bench rc=$?"; tail -n 1 gpurun_out/bench_r10.log | cut -c1-700
timeout 1500 python -m pytest tests/ -q -m gpu 2>&1 | tail -6
