set -x
N=${1:-8}
mkdir -p gpurun_out
nproc; free -g | head -2; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket" 
T="timeout 400 python -m torch.distributed.run --nproc-per-node $N --master-addr 127.0.0.1"
$T --master-port 29531 tests/gpu_zero_stress.py 200 > gpurun_out/zero_stress21_n$N.log 2>&1; echo "stress rc=$?"; tail -n 1 gpurun_out/zero_stress21_n$N.log | cut -c1-1500
LIBAI_B200_OVERLAP_GRAD_SYNC=0 $T --master-port 29532 bench.py --gpus $N --steps 8 --warmup 3 --no-e2e > gpurun_out/bench21_dp${N}_nooverlap.log 2>&1; echo "no-overlap rc=$?"; tail -n 1 gpurun_out/bench21_dp${N}_nooverlap.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','host_enqueue_ms_per_step','host_cpus','clocks','gpu_launches')})"
$T --master-port 29533 bench.py --gpus $N --steps 8 --warmup 3 --no-e2e > gpurun_out/bench21_dp${N}_overlap.log 2>&1; echo "overlap rc=$?"; tail -n 1 gpurun_out/bench21_dp${N}_overlap.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','host_enqueue_ms_per_step','host_cpus','clocks','gpu_launches')})"
