# round 2, run 15 (2 GPUs): after the 8-GPU findings — symmetric-memory teardown between layouts, the extras watchdog
# (injected failure must still give the JSON line and exit code 0), exactly-summable ZeRO check, CUDA graphs for blocks
# that take a key-padding mask (BERT) / RoPE buffers (Llama), ZeRO-2 with pooled gradient buckets + graphs
set -x
mkdir -p gpurun_out
export LIBAI_B200_SPIN_TIMEOUT_MS=30000
run() { timeout "$1" python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port "$2" "${@:3}"; }
run 600 29521 tests/gpu_comm_check.py --out gpurun_out/r2_15_comm_check_2gpu.json > gpurun_out/r2_15_comm_check_2gpu.log 2>&1
grep '"ok": false' gpurun_out/r2_15_comm_check_2gpu.log | cut -c1-600; tail -1 gpurun_out/r2_15_comm_check_2gpu.log
LIBAI_B200_BENCH_INJECT_EXTRA_FAIL=1 run 900 29522 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/r2_15_bench_2gpu_injected_extra_failure.json 2> gpurun_out/r2_15_inject.err
echo "injected-failure rc=$?"; tail -1 gpurun_out/r2_15_bench_2gpu_injected_extra_failure.json | cut -c1-1500
run 900 29523 bench.py --gpus 2 --steps 10 --warmup 4 > gpurun_out/r2_15_bench_2gpu.json 2> gpurun_out/r2_15_bench_2gpu.err
echo "bench rc=$?"; tail -1 gpurun_out/r2_15_bench_2gpu.json | cut -c1-2500
run 900 29524 bench.py --gpus 2 --zero 2 --steps 10 --warmup 4 --no-e2e --extras 0 --ref-same-box 0 > gpurun_out/r2_15_bench_2gpu_dp2_zero2_graphs.json 2> gpurun_out/r2_15_zero2.err
echo "zero2 rc=$?"; grep -i "cuda graphs" gpurun_out/r2_15_zero2.err | head -3 | cut -c1-300; tail -1 gpurun_out/r2_15_bench_2gpu_dp2_zero2_graphs.json | cut -c1-900
run 900 29525 bench.py --gpus 2 --model bert_large --layout tp2 --steps 8 --warmup 3 --no-e2e > gpurun_out/r2_15_bench_2gpu_bert_large_tp2.json 2> gpurun_out/r2_15_bert.err
echo "bert rc=$?"; grep -i "cuda graphs" gpurun_out/r2_15_bert.err | head -3 | cut -c1-300; tail -1 gpurun_out/r2_15_bench_2gpu_bert_large_tp2.json | cut -c1-900
run 900 29526 bench.py --gpus 2 --model llama7b --layers 8 --micro-batch 2 --zero 2 --steps 6 --warmup 3 --no-e2e > gpurun_out/r2_15_bench_2gpu_llama_8layers_dp2_zero2.json 2> gpurun_out/r2_15_llama.err
echo "llama rc=$?"; grep -i "cuda graphs" gpurun_out/r2_15_llama.err | head -3 | cut -c1-300; tail -1 gpurun_out/r2_15_bench_2gpu_llama_8layers_dp2_zero2.json | cut -c1-900
