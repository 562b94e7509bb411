set -x
mkdir -p gpurun_out
timeout 900 python tests/gpu_kernel_check.py --out gpurun_out/kc22.json > gpurun_out/kc22.log 2>&1; echo "kc rc=$?"
grep -E "SUMMARY|\"ok\": false" gpurun_out/kc22.log | cut -c1-600
grep -E "tflops" gpurun_out/kc22.log | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    if 'cublas_tflops' in d: print(d['name'], round(d['tflops']), 'vs cublas', round(d['cublas_tflops']))"
grep -E "linear_fwd speed|mlp fused|attention speed" gpurun_out/kc22.log | cut -c1-500
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r22.log 2>&1; echo "bench rc=$?"; tail -n 1 gpurun_out/bench_r22.log | cut -c1-700
timeout 1500 python -m pytest tests/ -q -m gpu 2>&1 | tail -n 4
