# round 2, run 24 (1 GPU): the other bench models (ViT-L images/s, GPT-2 large) and the reference recipe's dropout 0.1
set -x
mkdir -p gpurun_out
b() { name=$1; shift; timeout 600 python bench.py --steps 10 --warmup 4 --ref-same-box 0 --no-e2e "$@" > gpurun_out/r2_24_$name.json 2> gpurun_out/r2_24_$name.err; echo "$name rc=$?"; tail -1 gpurun_out/r2_24_$name.json | cut -c1-900; grep -i "cuda graphs\|Error" gpurun_out/r2_24_$name.err | tail -2 | cut -c1-250; }
b gpt2_dropout0.1 --dropout 0.1
b gpt2_dropout0 
b vit_l --model vit_l
b gpt2_large --model gpt2_large
