set -x
mkdir -p gpurun_out
T="timeout 500 python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1"
$T --master-port 29512 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/bench34_dp2.log 2>&1; echo "dp2 rc=$?"; tail -n 1 gpurun_out/bench34_dp2.log | cut -c1-1500
$T --master-port 29513 bench.py --gpus 2 --steps 6 --warmup 3 --tp 2 --no-e2e > gpurun_out/bench34_tp2.log 2>&1; echo "tp2 rc=$?"; tail -n 1 gpurun_out/bench34_tp2.log | cut -c1-330
