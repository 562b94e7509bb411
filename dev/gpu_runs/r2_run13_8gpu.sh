# round 2, run 13 (8 GPUs): the driver's N=8 command (dp8 + tp2xdp4 + tp2xpp2xdp2), fused collectives at 8 ranks,
# BERT-large tp2xdp4 and Llama-2-7B tp4xdp2 + ZeRO-2 (the other BASELINE.json configs)
set -x
mkdir -p gpurun_out
export LIBAI_B200_SPIN_TIMEOUT_MS=30000
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 8 --steps 10 --warmup 4 > gpurun_out/r2_bench_8gpu.json 2> gpurun_out/r2_bench_8gpu.err
tail -3 gpurun_out/r2_bench_8gpu.err | cut -c1-400; cat gpurun_out/r2_bench_8gpu.json | cut -c1-3500
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 29511 \
    tests/gpu_comm_check.py --out gpurun_out/r2_comm_check_8gpu.json > gpurun_out/r2_comm_check_8gpu.log 2>&1
grep '"ok": false' gpurun_out/r2_comm_check_8gpu.log | cut -c1-800; tail -1 gpurun_out/r2_comm_check_8gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus 8 --model bert_large --layout tp2 --steps 8 --warmup 3 --no-e2e > gpurun_out/r2_bench_8gpu_bert_large_tp2dp4.json 2> gpurun_out/r2_bench_8gpu_bert.err
tail -2 gpurun_out/r2_bench_8gpu_bert.err | cut -c1-300; cat gpurun_out/r2_bench_8gpu_bert_large_tp2dp4.json | cut -c1-900
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 29514 \
    bench.py --gpus 8 --model llama7b --layout tp4 --zero 2 --steps 5 --warmup 3 --no-e2e > gpurun_out/r2_bench_8gpu_llama7b_tp4dp2_zero2.json 2> gpurun_out/r2_bench_8gpu_llama.err
tail -2 gpurun_out/r2_bench_8gpu_llama.err | cut -c1-300; cat gpurun_out/r2_bench_8gpu_llama7b_tp4dp2_zero2.json | cut -c1-900
