set -x
mkdir -p gpurun_out
timeout 45 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2
timeout 50 python bench.py --steps 5 --warmup 3 --no-e2e > gpurun_out/r39_bench.log 2>&1; echo "bench rc=$?"; tail -n 1 gpurun_out/r39_bench.log | cut -c1-700
