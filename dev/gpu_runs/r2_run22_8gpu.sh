# round 2, run 22 (8 GPUs): the driver's N=8 command after the symmetric-memory teardown fix (dp8 + tp2xdp4 + tp2xpp2xdp2 +
# same-box PyTorch comparator), the fused collectives / ZeRO check at 8 ranks, Llama-2-7B tp4xdp2 + ZeRO-2 (CUDA graphs)
set -x
mkdir -p gpurun_out
export LIBAI_B200_SPIN_TIMEOUT_MS=30000
run() { timeout "$1" python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port "$2" "${@:3}"; }
run 900 29532 bench.py --gpus 8 --steps 10 --warmup 4 > gpurun_out/r2_22_bench_8gpu.json 2> gpurun_out/r2_22_bench_8gpu.err
echo "bench rc=$?"; tail -1 gpurun_out/r2_22_bench_8gpu.json | cut -c1-3000
run 400 29531 tests/gpu_comm_check.py --out gpurun_out/r2_22_comm_check_8gpu.json > gpurun_out/r2_22_comm_check_8gpu.log 2>&1
grep '"ok": false' gpurun_out/r2_22_comm_check_8gpu.log | cut -c1-800; tail -1 gpurun_out/r2_22_comm_check_8gpu.log
run 700 29534 bench.py --gpus 8 --model llama7b --layout tp4 --zero 2 --micro-batch 2 --steps 5 --warmup 3 --no-e2e > gpurun_out/r2_22_bench_8gpu_llama7b_tp4dp2_zero2.json 2> gpurun_out/r2_22_llama.err
echo "llama rc=$?"; grep -i "cuda graphs" gpurun_out/r2_22_llama.err | head -2 | cut -c1-250; tail -1 gpurun_out/r2_22_bench_8gpu_llama7b_tp4dp2_zero2.json | cut -c1-1000
run 500 29535 bench.py --gpus 8 --model bert_large --layout tp2 --steps 8 --warmup 3 --no-e2e > gpurun_out/r2_22_bench_8gpu_bert_large_tp2dp4.json 2> gpurun_out/r2_22_bert.err
echo "bert rc=$?"; tail -1 gpurun_out/r2_22_bench_8gpu_bert_large_tp2dp4.json | cut -c1-1000
