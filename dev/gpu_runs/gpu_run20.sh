set -x
mkdir -p gpurun_out
T="timeout 600 python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1"
$T --master-port 29511 tests/gpu_comm_check.py --out gpurun_out/comm20.json > gpurun_out/comm20.log 2>&1; echo "comm rc=$?"
grep -E "SUMMARY|\"ok\": false|overlapped|Traceback|Error" gpurun_out/comm20.log | cut -c1-600 | tail -n 12
$T --master-port 29512 bench.py --gpus 2 --steps 8 --warmup 3 --no-e2e > gpurun_out/bench20_dp2.log 2>&1; echo "dp2 rc=$?"; tail -n 1 gpurun_out/bench20_dp2.log | cut -c1-330
LIBAI_B200_OVERLAP_GRAD_SYNC=0 $T --master-port 29513 bench.py --gpus 2 --steps 8 --warmup 3 --no-e2e > gpurun_out/bench20_dp2_nooverlap.log 2>&1; echo "dp2 no-overlap rc=$?"; tail -n 1 gpurun_out/bench20_dp2_nooverlap.log | cut -c1-330
$T --master-port 29515 tests/gpu_zero_stress.py 300 > gpurun_out/zero_stress20.log 2>&1; echo "stress rc=$?"; tail -n 1 gpurun_out/zero_stress20.log | cut -c1-900
