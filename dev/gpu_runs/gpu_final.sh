set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2
timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -n 4 | cut -c1-300
timeout 600 python bench.py > gpurun_out/bench_final.log 2>&1; echo "bench rc=$?"; tail -n 1 gpurun_out/bench_final.log | cut -c1-1600
timeout 120 python bench.py --impl reference | tail -n 1
