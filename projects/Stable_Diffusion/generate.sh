#!/usr/bin/env bash
# sample with a fine-tuned checkpoint: bash projects/Stable_Diffusion/generate.sh <model_dir> "<prompt>" [lora_dir]
set -e
python - "$@" <<'PY'
import sys, torch
sys.path.insert(0, ".")
from projects.Stable_Diffusion.pipeline import StableDiffusionPipeline
model_dir, prompt = sys.argv[1], sys.argv[2]
dev = "cuda" if torch.cuda.is_available() else "cpu"
pipe = StableDiffusionPipeline.from_pretrained(model_dir).to(dev, torch.bfloat16 if dev == "cuda" else torch.float32)
if len(sys.argv) > 3:
    pipe.load_lora(sys.argv[3])
pipe(prompt)[0].save("sd_sample.png")
print("wrote sd_sample.png")
PY
