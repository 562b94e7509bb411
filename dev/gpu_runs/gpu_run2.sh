set -x
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
tail -n 3 gpurun_out/smoke.log; tail -n 2 gpurun_out/bench.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 2 --no-e2e > gpurun_out/bench_ncu.log 2>&1; echo "ncu rc=$?"
