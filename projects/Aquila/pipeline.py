"""Aquila text generation pipeline (reference projects/Aquila/pipeline.py)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))

from libai_b200.utils import distributed as dist  # noqa: E402
from projects.common.llm_pipeline import CausalLMPipeline  # noqa: E402
from projects.Aquila.utils.aquila_loader import AquilaLoaderHuggerFace, AquilaLoaderLiBai  # noqa: E402


class TextGenerationPipeline(CausalLMPipeline):
    hf_loader = AquilaLoaderHuggerFace
    libai_loader = AquilaLoaderLiBai


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--config_file", default="projects/Aquila/configs/aquila_config.py")
    ap.add_argument("--model_path", default=None)
    ap.add_argument("--mode", default="huggingface", choices=["huggingface", "libai", "random"])
    ap.add_argument("--tensor_parallel", type=int, default=1)
    ap.add_argument("--pipeline_parallel", type=int, default=1)
    ap.add_argument("--prompt", default="Give three tips for staying healthy.")
    args = ap.parse_args()
    pipeline = TextGenerationPipeline(args.config_file, data_parallel=1, tensor_parallel=args.tensor_parallel,
                                      pipeline_parallel=args.pipeline_parallel,
                                      pipeline_num_layers=32 if args.pipeline_parallel > 1 else None,
                                      model_path=args.model_path, mode=args.mode)
    out = pipeline([args.prompt])
    if dist.is_main_process():
        print(out)
