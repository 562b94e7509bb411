set -x
mkdir -p gpurun_out
timeout 600 python tests/gpu_kernel_check.py --only "embedding,cross" --out gpurun_out/kc33.json > gpurun_out/kc33.log 2>&1; echo "kc rc=$?"
grep -E "SUMMARY|\"ok\": false|embedding|Traceback|Error" gpurun_out/kc33.log | cut -c1-700
timeout 900 python -m pytest tests/test_gpu.py -q -x -k "native_training or cuda_graph or smoke" 2>&1 | tail -n 4 | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r33.log 2>&1; echo "bench rc=$?"; tail -n 1 gpurun_out/bench_r33.log | cut -c1-1600
