# round 2, run 27 (2 GPUs): host time of ZeRO-2 with 7B-width buckets after the persistent reduce-scatter outputs; cProfile
set -x
mkdir -p gpurun_out
run() { timeout "$1" python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port "$2" "${@:3}"; }
run 600 29551 bench.py --gpus 2 --model llama7b --layers 8 --micro-batch 2 --zero 2 --steps 6 --warmup 3 --no-e2e > gpurun_out/r2_27_bench_2gpu_llama_8layers_dp2_zero2.json 2> gpurun_out/r2_27_llama.err
echo "llama rc=$?"; tail -1 gpurun_out/r2_27_bench_2gpu_llama_8layers_dp2_zero2.json | cut -c1-900
run 600 29552 dev/profile_host.py --layout dp --steps 4 --model llama7b --layers 8 --micro-batch 2 --zero 2 > gpurun_out/r2_27_host_profile.log 2>&1
echo "profile rc=$?"; head -60 gpurun_out/host_profile_dp_rank0.txt | cut -c1-180
