set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > gpurun_out/clocks17.csv &
SMI=$!
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r17.log 2>&1; echo "bench rc=$?"; tail -n 1 gpurun_out/bench_r17.log | cut -c1-900
kill $SMI
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1400 -c 1500 --csv --log-file gpurun_out/launches_r17.csv python bench.py --steps 1 --warmup 1 --no-e2e > gpurun_out/bench_ncu17.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 3 -c 2 -o gpurun_out/prof17_gemm python tests/gpu_kernel_check.py --only "linear_fwd speed" --out gpurun_out/tmp.json > gpurun_out/ncu17_gemm.log 2>&1; echo "ncu gemm rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 4 -c 1 -o gpurun_out/prof17_attn_fwd python tests/gpu_kernel_check.py --only "attention speed" --out gpurun_out/tmp.json > gpurun_out/ncu17_attn_fwd.log 2>&1; echo "ncu attn fwd rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_bwd -s 4 -c 1 -o gpurun_out/prof17_attn_bwd python tests/gpu_kernel_check.py --only "attention speed" --out gpurun_out/tmp.json > gpurun_out/ncu17_attn_bwd.log 2>&1; echo "ncu attn bwd rc=$?"
ls -la gpurun_out/prof17*.ncu-rep
