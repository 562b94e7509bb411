"""Blank-infilling generation with GLM (reference projects/GLM/infer_glm.py).

    bash tools/infer.sh projects/GLM/infer_glm.py 4 --model /path/to/glm-10b-chinese --tp 2 --pp 2
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))

import torch  # noqa: E402

from libai_b200.config import LazyConfig  # noqa: E402
from libai_b200.utils import distributed as dist  # noqa: E402
from projects.GLM.modeling_glm import GLMForConditionalGeneration  # noqa: E402
from projects.GLM.tokenizer.glm_tokenizer import GLMChineseTokenzier, GLMGPT2Tokenizer  # noqa: E402
from projects.GLM.utils.glm_loader import GLMLoaderHuggerFace  # noqa: E402

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", required=True)
    ap.add_argument("--tp", type=int, default=1)
    ap.add_argument("--pp", type=int, default=1)
    ap.add_argument("--prompt", default="Ng is an adjunct professor at [MASK] (formerly associate professor).")
    args = ap.parse_args()
    cfg = LazyConfig.load("projects/GLM/configs/glm_inference.py")
    train = LazyConfig.load("configs/common/train.py").train
    train.dist.tensor_parallel_size, train.dist.pipeline_parallel_size = args.tp, args.pp
    train.dist.pipeline_num_layers = cfg.cfg.num_layers
    dist.setup_dist_util(train.dist)
    spm_file = os.path.join(args.model, "cog-pretrain.model")
    tokenizer = GLMChineseTokenzier(spm_file) if os.path.exists(spm_file) else GLMGPT2Tokenizer(
        os.path.join(args.model, "vocab.json"), os.path.join(args.model, "merges.txt"))
    model = GLMLoaderHuggerFace(GLMForConditionalGeneration, cfg.cfg, args.model).load().eval()
    if torch.cuda.is_available():
        model = model.cuda().bfloat16()
    inputs = tokenizer(args.prompt)
    inputs = tokenizer.build_inputs_for_generation(inputs, max_gen_length=64)
    dev = next(model.parameters()).device
    out = model.generate(inputs["input_ids"].to(dev), position_ids=inputs["position_ids"].to(dev),
                         generation_attention_mask=inputs["generation_attention_mask"].to(dev), max_length=inputs["input_ids"].shape[1] + 63,
                         eos_token_id=tokenizer.eop_token_id)
    if dist.is_main_process():
        print(tokenizer.decode(out[0].tolist()))
