set -x
timeout 300 python tests/gpu_kernel_check.py --only norm,layernorm,cross,attention,gemm\ L2 --out gpurun_out/kc2.json > gpurun_out/kc2.log 2>&1; echo "kc2 rc=$?"
grep -E "FAIL|SUMMARY|speed|cross|L2 (1024|4096)" gpurun_out/kc2.log | cut -c1-400
timeout 600 python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/gpu_comm_check.py > gpurun_out/comm.log 2>&1; echo "comm rc=$?"
tail -n 25 gpurun_out/comm.log | cut -c1-600
timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/bench1.log 2>&1; echo "bench1 rc=$?"; tail -n 1 gpurun_out/bench1.log | cut -c1-900
timeout 600 python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/bench2.log 2>&1; echo "bench2 rc=$?"; tail -n 1 gpurun_out/bench2.log | cut -c1-900
timeout 600 python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 8 --warmup 3 --zero 1 --no-e2e > gpurun_out/bench2z.log 2>&1; echo "bench2z rc=$?"; tail -n 1 gpurun_out/bench2z.log | cut -c1-600
