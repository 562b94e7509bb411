set -x
timeout 400 python tests/gpu_kernel_check.py --only gemm --out gpurun_out/kc_gemm_bulk.json > gpurun_out/kc_gemm_bulk.log 2>&1; echo "gemm(bulk) rc=$?"
grep -E "FAIL|\"ok\": false|SUMMARY" gpurun_out/kc_gemm_bulk.log | cut -c1-300
grep -E "tflops" gpurun_out/kc_gemm_bulk.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['name'], round(d['tflops']), 'vs cublas', round(d['cublas_tflops']))"
LIBAI_B200_WGRAD_BULK=0 timeout 300 python tests/gpu_kernel_check.py --only "gemm L2" --out gpurun_out/kc_gemm_atomic.json > gpurun_out/kc_gemm_atomic.log 2>&1; echo "gemm(atomic) rc=$?"
grep -E "tflops" gpurun_out/kc_gemm_atomic.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('ATOMIC', d['name'], round(d['tflops']), 'vs cublas', round(d['cublas_tflops']))"
timeout 300 python tests/gpu_kernel_check.py --only attention --out gpurun_out/kc_attn_v2.json > gpurun_out/kc_attn_v2.log 2>&1; echo "attn v2 rc=$?"
grep -E "attention" gpurun_out/kc_attn_v2.log | cut -c1-330
LIBAI_B200_ATTN_FWD=1 timeout 300 python tests/gpu_kernel_check.py --only "attention speed" --out gpurun_out/kc_attn_v1.json > gpurun_out/kc_attn_v1.log 2>&1
grep -E "attention speed" gpurun_out/kc_attn_v1.log | cut -c1-330
timeout 300 python tests/gpu_kernel_check.py --only norm,layernorm,cross,elementwise,fused,rope,linear --out gpurun_out/kc_misc2.json > gpurun_out/kc_misc2.log 2>&1; grep -E "SUMMARY" gpurun_out/kc_misc2.log
timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_r4.log 2>&1; echo "bench rc=$?"; tail -n 1 gpurun_out/bench_r4.log | cut -c1-1200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2500 -c 1300 --csv --log-file gpurun_out/launches_r4.csv python bench.py --steps 2 --warmup 2 --no-e2e > gpurun_out/bench_ncu4.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 40 -c 3 -o gpurun_out/prof_gemm python tests/gpu_kernel_check.py --only "gemm L0 8192x4096x1024" --out gpurun_out/tmp.json > gpurun_out/ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 10 -c 2 -o gpurun_out/prof_attn python tests/gpu_kernel_check.py --only "attention speed" --out gpurun_out/tmp.json > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn rc=$?"
ls -la gpurun_out/*.ncu-rep
