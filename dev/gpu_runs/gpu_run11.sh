set -x
mkdir -p gpurun_out
T="timeout 600 python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1"
$T --master-port 29511 tests/gpu_comm_check.py --out gpurun_out/comm11.json > gpurun_out/comm11.log 2>&1; echo "comm rc=$?"
grep -E "SUMMARY|\"ok\": false|speed|ms" gpurun_out/comm11.log | cut -c1-500 | tail -n 20
$T --master-port 29512 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/bench11_dp2.log 2>&1; echo "dp2 rc=$?"; tail -n 1 gpurun_out/bench11_dp2.log | cut -c1-400
$T --master-port 29513 bench.py --gpus 2 --steps 8 --warmup 3 --tp 2 > gpurun_out/bench11_tp2.log 2>&1; echo "tp2 rc=$?"; tail -n 1 gpurun_out/bench11_tp2.log | cut -c1-400
$T --master-port 29514 bench.py --gpus 2 --steps 8 --warmup 3 --pp 2 --acc 4 --micro-batch 2 > gpurun_out/bench11_pp2.log 2>&1; echo "pp2 rc=$?"; tail -n 1 gpurun_out/bench11_pp2.log | cut -c1-400
timeout 900 python -m pytest tests/ -q -m gpu -k "nvlink" 2>&1 | tail -n 5
