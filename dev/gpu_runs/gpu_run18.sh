# usage: bash dev/gpu_runs/gpu_run18.sh N  — N-GPU validation: fused collectives at world=N, then the scaling bench layouts
set -x
N=${1:-4}
mkdir -p gpurun_out
T="timeout 600 python -m torch.distributed.run --nproc-per-node $N --master-addr 127.0.0.1"
$T --master-port 29521 tests/gpu_comm_check.py --out gpurun_out/comm18_n$N.json > gpurun_out/comm18_n$N.log 2>&1; echo "comm rc=$?"
grep -E "SUMMARY|\"ok\": false|ms" gpurun_out/comm18_n$N.log | cut -c1-500 | tail -n 12
$T --master-port 29522 bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/bench18_dp$N.log 2>&1; echo "dp$N rc=$?"; tail -n 1 gpurun_out/bench18_dp$N.log | cut -c1-700
$T --master-port 29523 bench.py --gpus $N --steps 6 --warmup 3 --tp 2 --no-e2e > gpurun_out/bench18_tp2_n$N.log 2>&1; echo "tp2 rc=$?"; tail -n 1 gpurun_out/bench18_tp2_n$N.log | cut -c1-300

