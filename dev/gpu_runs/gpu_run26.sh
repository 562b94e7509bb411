set -x
mkdir -p gpurun_out
N=${1:-4}
T="timeout 500 python -m torch.distributed.run --nproc-per-node $N --master-addr 127.0.0.1"
$T --master-port 29512 bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/bench26_dp$N.log 2>&1; echo "dp$N rc=$?"; tail -n 1 gpurun_out/bench26_dp$N.log | cut -c1-1500
