set -x
mkdir -p gpurun_out
timeout 300 python dev/gpu_runs/wgrad_sweep.py 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    p = l.split(' ', 2)
    try: d = json.loads(p[2])
    except Exception: continue
    print(p[0], p[1], {k: d[k] for k in ('bn0_s0', 'bn0_s1', 'bn0_s2', 'bn256_s0', 'cublas_ms')})"
timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e > gpurun_out/bench_r29.log 2>&1; echo "bench rc=$?"; tail -n 1 gpurun_out/bench_r29.log | cut -c1-330
LIBAI_B200_WGRAD_RMW=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e > gpurun_out/bench_r29_rmw.log 2>&1; echo "bench rmw rc=$?"; tail -n 1 gpurun_out/bench_r29_rmw.log | cut -c1-330
