set -x
mkdir -p gpurun_out
T="timeout 600 python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1"
$T --master-port 29511 tests/gpu_comm_check.py --out gpurun_out/comm16.json > gpurun_out/comm16.log 2>&1; echo "comm rc=$?"
grep -E "SUMMARY|\"ok\": false|ms" gpurun_out/comm16.log | cut -c1-400 | tail -n 12
$T --master-port 29513 bench.py --gpus 2 --steps 8 --warmup 3 --tp 2 --no-e2e > gpurun_out/bench16_tp2.log 2>&1; echo "tp2 rc=$?"; tail -n 1 gpurun_out/bench16_tp2.log | cut -c1-300
