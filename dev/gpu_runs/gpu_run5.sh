set -x
mkdir -p gpurun_out
timeout 600 python tests/gpu_kernel_check.py --out gpurun_out/kc5.json > gpurun_out/kc5.log 2>&1; echo "kc rc=$?"
grep -E "SUMMARY|\"ok\": false" gpurun_out/kc5.log | cut -c1-400
grep -E "tflops" gpurun_out/kc5.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    if 'cublas_tflops' in d: print(d['name'], round(d['tflops']), 'vs cublas', round(d['cublas_tflops']))"
grep -E "linear_fwd speed|mlp fused|attention speed|rope_qkv" gpurun_out/kc5.log | cut -c1-400
timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_r5.log 2>&1; echo "bench rc=$?"; tail -n 1 gpurun_out/bench_r5.log | cut -c1-1500
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1600 -c 1700 --csv --log-file gpurun_out/launches_r5.csv python bench.py --steps 1 --warmup 1 --no-e2e > gpurun_out/bench_ncu5.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 3 -c 2 -o gpurun_out/prof_gemm python tests/gpu_kernel_check.py --only "linear_fwd speed" --out gpurun_out/tmp.json > gpurun_out/ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 8 -c 3 -o gpurun_out/prof_attn python tests/gpu_kernel_check.py --only "attention speed" --out gpurun_out/tmp.json > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn rc=$?"
ls -la gpurun_out/*.ncu-rep
