# round 2, run 38 (2 GPUs): fused gate/up under tensor parallelism (one AG->GEMM / one GEMM->RS instead of two each):
# parity, then an 8-layer Llama-7B-width model tp2 with and without
set -x
mkdir -p gpurun_out
export LIBAI_B200_SPIN_TIMEOUT_MS=20000
run() { timeout "$1" python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port "$2" "${@:3}"; }
run 150 29581 tests/gpu_gated_mlp_check.py > gpurun_out/r2_38_gated_mlp_tp2.json 2> gpurun_out/r2_38_gated_mlp_tp2.err; echo "parity rc=$?"; tail -1 gpurun_out/r2_38_gated_mlp_tp2.json | cut -c1-600; grep -i "error" gpurun_out/r2_38_gated_mlp_tp2.err | tail -2 | cut -c1-300
run 300 29582 bench.py --gpus 2 --model llama7b --layers 8 --layout tp2 --micro-batch 2 --steps 5 --warmup 3 --no-e2e > gpurun_out/r2_38_bench_2gpu_llama_8layers_tp2_fused_gate_up.json 2> gpurun_out/r2_38_a.err; echo "fused rc=$?"; tail -1 gpurun_out/r2_38_bench_2gpu_llama_8layers_tp2_fused_gate_up.json | cut -c1-330
LIBAI_B200_FUSED_GATE_UP=0 run 300 29583 bench.py --gpus 2 --model llama7b --layers 8 --layout tp2 --micro-batch 2 --steps 5 --warmup 3 --no-e2e > gpurun_out/r2_38_bench_2gpu_llama_8layers_tp2_two_gemms.json 2> gpurun_out/r2_38_b.err; echo "two-gemm rc=$?"; tail -1 gpurun_out/r2_38_bench_2gpu_llama_8layers_tp2_two_gemms.json | cut -c1-330
